"""ctypes binding of oracle/libflac_oracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (flac_codec_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libflac_oracle.so")


def build(force: bool = False) -> str:
    """Compile the oracle with the recipe in oracle/Makefile (gcc only)."""
    src = os.path.join(_HERE, "flac_oracle.c")
    deps = [p for p in (src, os.path.join(_HERE, "flac_oracle.h"), os.path.join(_HERE, "Makefile")) if os.path.exists(p)]
    if force or not os.path.exists(_LIB_PATH) or any(os.path.getmtime(p) > os.path.getmtime(_LIB_PATH) for p in deps):
        subprocess.check_call(["make", "-C", _HERE, "-s"], env={**os.environ, "CC": "gcc"})
    return _LIB_PATH


class Options(C.Structure):
    _fields_ = [
        ("block_size", C.c_uint16),
        ("max_lpc_order", C.c_uint8),
        ("max_partition_order", C.c_uint8),
        ("mid_side", C.c_uint8),
        ("exhaustive_channel_correlation", C.c_uint8),
        ("window_kind", C.c_uint8),
        ("tukey_p", C.c_float),
        ("seektable_kind", C.c_uint8),
        ("seektable_n", C.c_uint32),
        ("padding", C.c_int32),
    ]


class SubframeInfo(C.Structure):
    _fields_ = [
        ("type", C.c_int32),
        ("order", C.c_int32),
        ("wasted", C.c_int32),
        ("bps", C.c_int32),
        ("precision", C.c_int32),
        ("shift", C.c_int32),
        ("coefs", C.c_int32 * 32),
        ("coding_method", C.c_int32),
        ("partition_order", C.c_int32),
        ("rice", C.c_uint8 * 64),
        ("kind", C.c_uint8 * 64),
        ("bits", C.c_uint64),
    ]


class FrameInfo(C.Structure):
    _fields_ = [
        ("channel_assignment", C.c_int32),
        ("channels", C.c_int32),
        ("frame_bytes", C.c_uint32),
        ("sub", SubframeInfo * 8),
    ]


class Streaminfo(C.Structure):
    _fields_ = [
        ("min_block_size", C.c_uint16),
        ("max_block_size", C.c_uint16),
        ("min_frame_size", C.c_uint32),
        ("max_frame_size", C.c_uint32),
        ("sample_rate", C.c_uint32),
        ("channels", C.c_uint8),
        ("bps", C.c_uint8),
        ("total_samples", C.c_uint64),
        ("md5", C.c_uint8 * 16),
        ("frames_start", C.c_uint64),
    ]


class FrameHeader(C.Structure):
    _fields_ = [
        ("block_size", C.c_uint32),
        ("sample_rate", C.c_uint32),
        ("bps", C.c_uint32),
        ("channels", C.c_uint32),
        ("channel_assignment", C.c_uint32),
        ("blocking_strategy", C.c_uint32),
        ("frame_number", C.c_uint64),
        ("header_bytes", C.c_uint32),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        u8p, i32p, f64p, u32p, u64p = (C.POINTER(t) for t in (C.c_uint8, C.c_int32, C.c_double, C.c_uint32, C.c_uint64))
        L.fo_encoder_new.restype = C.c_void_p
        L.fo_encoder_free.argtypes = [C.c_void_p]
        L.fo_encode_frame.restype = C.c_int64
        L.fo_encode_frame.argtypes = [C.c_void_p, C.POINTER(Options), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64,
                                      C.POINTER(i32p), C.c_uint32, C.c_int, u8p, C.c_size_t, C.POINTER(FrameInfo)]
        L.fo_encode_stream.restype = C.c_int64
        L.fo_encode_stream.argtypes = [C.POINTER(Options), C.c_uint32, C.c_uint32, C.c_uint32, i32p, C.c_uint64, C.c_int,
                                       C.c_int, u8p, C.c_size_t, u32p, C.c_size_t, u64p]
        L.fo_encode_frames_only.restype = C.c_int64
        L.fo_encode_frames_only.argtypes = [C.POINTER(Options), C.c_uint32, C.c_uint32, C.c_uint32, i32p, C.c_uint64,
                                            C.c_uint64, C.c_int, u8p, C.c_size_t, u32p, C.c_size_t, u64p,
                                            C.POINTER(FrameInfo)]
        L.fo_read_streaminfo.restype = C.c_int
        L.fo_read_streaminfo.argtypes = [u8p, C.c_size_t, C.POINTER(Streaminfo)]
        L.fo_decode_frame.restype = C.c_int64
        L.fo_decode_frame.argtypes = [u8p, C.c_size_t, C.POINTER(Streaminfo), C.c_uint64, i32p, C.c_size_t,
                                      C.POINTER(FrameHeader)]
        L.fo_decode_stream.restype = C.c_int64
        L.fo_decode_stream.argtypes = [u8p, C.c_size_t, i32p, C.c_size_t, C.POINTER(Streaminfo), u8p]
        L.fo_decode_frames_mt.restype = C.c_int64
        L.fo_decode_frames_mt.argtypes = [u8p, u64p, C.c_uint64, C.POINTER(Streaminfo), C.c_int, i32p, C.c_size_t]
        L.fo_autocorrelate.argtypes = [f64p, C.c_uint32, C.c_uint32, f64p]
        L.fo_lp_coefficients.argtypes = [f64p, C.c_uint32, f64p, f64p]
        L.fo_subframe_bits_by_order.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, f64p, C.c_uint32, f64p]
        L.fo_quantize.argtypes = [C.c_uint32, f64p, C.c_uint32, i32p, u32p]
        L.fo_lpc_residuals.argtypes = [C.c_uint32, C.c_uint32, i32p, i32p, C.c_uint32, i32p]
        L.fo_predict.argtypes = [C.POINTER(C.c_int64), C.c_uint32, C.c_uint32, i32p, C.c_uint32]
        L.fo_predict.restype = None
        L.fo_window.argtypes = [C.POINTER(Options), C.c_uint32, f64p]
        L.fo_window.restype = None
        L.fo_rice_parameter_f64.restype = C.c_uint32
        L.fo_rice_parameter_f64.argtypes = [C.c_uint64, C.c_uint32]
        L.fo_crc8.restype = C.c_uint8
        L.fo_crc8.argtypes = [u8p, C.c_size_t]
        L.fo_crc16.restype = C.c_uint16
        L.fo_crc16.argtypes = [u8p, C.c_size_t]
        L.fo_md5.argtypes = [u8p, C.c_size_t, u8p]
        L.fo_md5.restype = None
        L.fo_write_frame_number.argtypes = [C.c_uint64, u8p]
        L.fo_read_frame_number.argtypes = [u8p, C.c_size_t, u64p]
        L.fo_bytes_to_samples.argtypes = [u8p, C.c_size_t, C.c_uint32, C.c_int, i32p]
        L.fo_bytes_to_samples.restype = None
        L.fo_samples_to_bytes.argtypes = [i32p, C.c_size_t, C.c_uint32, C.c_int, u8p]
        L.fo_samples_to_bytes.restype = None
        for name in ("fo_options_default", "fo_options_fast", "fo_options_best"):
            getattr(L, name).argtypes = [C.POINTER(Options)]
            getattr(L, name).restype = None
        _lib = L
    return _lib


class OracleError(Exception):
    def __init__(self, code: int):
        super().__init__(f"oracle error code {code}")
        self.code = code


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def options(preset: str = "default", **kw) -> Options:
    o = Options()
    getattr(lib(), f"fo_options_{preset}")(C.byref(o))
    for k, v in kw.items():
        if k == "max_lpc_order" and v is None:
            v = 0
        if k == "padding" and v is None:
            v = -1
        setattr(o, k, v)
    return o


def encode_frame(opt: Options, sample_rate: int, bps: int, planar: np.ndarray, frame_number: int = 0,
                 subset: bool = False, want_info: bool = False):
    """planar: int32 [channels, n]. Returns bytes (and FrameInfo)."""
    L = lib()
    planar = np.ascontiguousarray(planar, dtype=np.int32)
    ch, n = planar.shape
    ptrs = (C.POINTER(C.c_int32) * ch)(*[_p(planar[c], C.c_int32) for c in range(ch)])
    out = np.zeros(n * ch * 5 + 256, dtype=np.uint8)
    info = FrameInfo()
    e = L.fo_encoder_new()
    try:
        r = L.fo_encode_frame(e, C.byref(opt), sample_rate, bps, ch, frame_number, ptrs, n, int(subset),
                              _p(out, C.c_uint8), out.size, C.byref(info))
    finally:
        L.fo_encoder_free(e)
    if r < 0:
        raise OracleError(-r)
    b = out[:r].tobytes()
    return (b, info) if want_info else b


def encode_stream(opt: Options, sample_rate: int, bps: int, channels: int, interleaved: np.ndarray,
                  total_known: bool = True, nthreads: int = 1):
    """Returns (flac_bytes, frame_sizes)."""
    L = lib()
    x = np.ascontiguousarray(interleaved, dtype=np.int32).reshape(-1)
    n_pcm = x.size // channels
    nf = (n_pcm + opt.block_size - 1) // opt.block_size
    cap = x.size * 5 + 256 * (nf + 1) + 8192 + 18 * (nf + 1) + max(opt.padding, 0)
    out = np.zeros(cap, dtype=np.uint8)
    sizes = np.zeros(max(nf, 1), dtype=np.uint32)
    nfo = C.c_uint64(0)
    r = L.fo_encode_stream(C.byref(opt), sample_rate, bps, channels, _p(x, C.c_int32), n_pcm, int(total_known),
                           nthreads, _p(out, C.c_uint8), cap, _p(sizes, C.c_uint32), sizes.size, C.byref(nfo))
    if r < 0:
        raise OracleError(-r)
    return out[:r].tobytes(), sizes[: nfo.value].copy()


def encode_frames_only(opt: Options, sample_rate: int, bps: int, channels: int, interleaved: np.ndarray,
                       first_frame_number: int = 0, nthreads: int = 1, want_infos: bool = False):
    """Frames without container. Returns (bytes, frame_sizes[, infos])."""
    L = lib()
    x = np.ascontiguousarray(interleaved, dtype=np.int32).reshape(-1)
    n_pcm = x.size // channels
    nf = (n_pcm + opt.block_size - 1) // opt.block_size
    cap = x.size * 5 + 256 * (nf + 1)
    out = np.zeros(cap, dtype=np.uint8)
    sizes = np.zeros(max(nf, 1), dtype=np.uint32)
    infos = (FrameInfo * max(nf, 1))() if want_infos else None
    nfo = C.c_uint64(0)
    r = L.fo_encode_frames_only(C.byref(opt), sample_rate, bps, channels, _p(x, C.c_int32), n_pcm, first_frame_number,
                                nthreads, _p(out, C.c_uint8), cap, _p(sizes, C.c_uint32), sizes.size, C.byref(nfo), infos)
    if r < 0:
        raise OracleError(-r)
    res = (out[:r].tobytes(), sizes[: nfo.value].copy())
    return res + (infos,) if want_infos else res


def read_streaminfo(flac: bytes) -> Streaminfo:
    L = lib()
    a = np.frombuffer(flac, dtype=np.uint8)
    si = Streaminfo()
    rc = L.fo_read_streaminfo(_p(a, C.c_uint8), a.size, C.byref(si))
    if rc:
        raise OracleError(rc)
    return si


def decode_stream(flac: bytes, want_md5: bool = False):
    """Returns (interleaved int32 samples, Streaminfo[, md5 bytes])."""
    L = lib()
    a = np.frombuffer(flac, dtype=np.uint8)
    si = read_streaminfo(flac)
    if si.total_samples:
        cap = si.total_samples * si.channels
    else:
        cap = max(len(flac) * 64, 1 << 16)
    out = np.zeros(cap, dtype=np.int32)
    md5 = np.zeros(16, dtype=np.uint8)
    r = L.fo_decode_stream(_p(a, C.c_uint8), a.size, _p(out, C.c_int32), out.size, C.byref(si), _p(md5, C.c_uint8))
    if r < 0:
        raise OracleError(-r)
    res = (out[:r].copy(), si)
    return res + (md5.tobytes(),) if want_md5 else res


def decode_stream_ex(flac: bytes):
    """The serial reader's full verdict: (error code or 0, frames delivered before the error, interleaved samples
    delivered, the samples).  Decoder::read_frame hands out every good frame before it fails (src/decode.rs:1388)."""
    L = lib()
    L.fo_decode_stream_ex.restype = C.c_int64
    L.fo_decode_stream_ex.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(Streaminfo), C.c_void_p,
                                      C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    a = np.frombuffer(flac, dtype=np.uint8)
    si = read_streaminfo(flac)
    # (room for frames that overshoot an understated total: the serial reader delivers them before it fails)
    cap = (si.total_samples + 4 * 65536) * si.channels if si.total_samples else max(len(flac) * 64, 1 << 16)
    out = np.zeros(cap, dtype=np.int32)
    nf, ns = C.c_uint64(0), C.c_uint64(0)
    r = L.fo_decode_stream_ex(a.ctypes.data, a.size, out.ctypes.data, out.size, C.byref(si), None, C.byref(nf), C.byref(ns))
    return (int(-r) if r < 0 else 0), nf.value, ns.value, out[: ns.value]


def decode_frame(data: bytes, si: Streaminfo | None = None, remaining: int = 0):
    """Returns (planar int32 [channels, block], FrameHeader, bytes_consumed)."""
    L = lib()
    a = np.frombuffer(data, dtype=np.uint8)
    out = np.zeros(65536 * 8, dtype=np.int32)
    h = FrameHeader()
    r = L.fo_decode_frame(_p(a, C.c_uint8), a.size, C.byref(si) if si is not None else None, remaining,
                          _p(out, C.c_int32), out.size, C.byref(h))
    if r < 0:
        raise OracleError(-r)
    return out[: h.channels * h.block_size].reshape(h.channels, h.block_size).copy(), h, r


def decode_frames_mt(frames: bytes, offsets: np.ndarray, si: Streaminfo, nthreads: int, total_samples: int):
    L = lib()
    a = np.frombuffer(frames, dtype=np.uint8)
    offs = np.ascontiguousarray(offsets, dtype=np.uint64)
    nfr = offs.size - 1
    out = np.zeros(nfr * si.max_block_size * si.channels, dtype=np.int32)
    r = L.fo_decode_frames_mt(_p(a, C.c_uint8), _p(offs, C.c_uint64), nfr, C.byref(si), nthreads, _p(out, C.c_int32), out.size)
    if r < 0:
        raise OracleError(-r)
    return out[: total_samples * si.channels]


def libm(fn: int, x: np.ndarray) -> np.ndarray:
    """glibc's log (fn 0) / log2 (fn 1), elementwise (numpy's own log is a SIMD implementation, not the C library's)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    L = lib()
    L.fo_libm.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
    L.fo_libm.restype = None
    L.fo_libm(fn, x.ctypes.data, out.ctypes.data, x.size)
    return out


def md5(b: bytes) -> bytes:
    a = np.frombuffer(b, dtype=np.uint8)
    o = np.zeros(16, dtype=np.uint8)
    lib().fo_md5(_p(a, C.c_uint8), a.size, _p(o, C.c_uint8))
    return o.tobytes()


def crc8(b: bytes) -> int:
    a = np.frombuffer(b, dtype=np.uint8)
    return lib().fo_crc8(_p(a, C.c_uint8), a.size)


def crc16(b: bytes) -> int:
    a = np.frombuffer(b, dtype=np.uint8)
    return lib().fo_crc16(_p(a, C.c_uint8), a.size)


def bytes_to_samples(b: bytes, bytes_per_sample: int, big_endian: bool = False) -> np.ndarray:
    a = np.frombuffer(b, dtype=np.uint8)
    n = a.size // bytes_per_sample
    out = np.zeros(n, dtype=np.int32)
    lib().fo_bytes_to_samples(_p(a, C.c_uint8), n, bytes_per_sample, int(big_endian), _p(out, C.c_int32))
    return out


def samples_to_bytes(s: np.ndarray, bytes_per_sample: int, big_endian: bool = False) -> bytes:
    s = np.ascontiguousarray(s, dtype=np.int32).reshape(-1)
    out = np.zeros(s.size * bytes_per_sample, dtype=np.uint8)
    lib().fo_samples_to_bytes(_p(s, C.c_int32), s.size, bytes_per_sample, int(big_endian), _p(out, C.c_uint8))
    return out.tobytes()
