"""ctypes binding of include/flacb200.h (libflacb200.so).

This is the only way Python reaches the engine: through the same C ABI a Rust/C++ host would bind.
There is no CPU fallback -- loading fails loudly when the library or the GPU is missing.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FLACB200_LIB") or os.path.join(_HERE, "libflacb200.so")   # FLACB200_LIB: A/B builds of the same library

HOST, DEVICE = 0, 1
PCM_BYTES_LE, PCM_BYTES_BE, PCM_I32_INTERLEAVED, PCM_I32_PLANAR = 0, 1, 2, 3


class Options(C.Structure):
    _fields_ = [
        ("block_size", C.c_uint16),
        ("max_lpc_order", C.c_uint8),
        ("max_partition_order", C.c_uint8),
        ("mid_side", C.c_uint8),
        ("exhaustive_channel_correlation", C.c_uint8),
        ("window_kind", C.c_uint8),
        ("reserved0", C.c_uint8),
        ("tukey_p", C.c_float),
    ]


class StreamParams(C.Structure):
    _fields_ = [
        ("sample_rate", C.c_uint32),
        ("bits_per_sample", C.c_uint32),
        ("channels", C.c_uint32),
        ("subset", C.c_uint32),
        ("max_block_size", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


class Segment(C.Structure):
    _fields_ = [("pcm_offset", C.c_uint64), ("n_pcm_frames", C.c_uint64), ("first_frame_number", C.c_uint64)]


class DecodeSegment(C.Structure):
    _fields_ = [("byte_offset", C.c_uint64), ("byte_length", C.c_uint64), ("pcm_offset", C.c_uint64),
                ("n_pcm_frames", C.c_uint64)]


class SubframeInfo(C.Structure):
    _fields_ = [
        ("type", C.c_int32), ("order", C.c_int32), ("wasted", C.c_int32), ("bps", C.c_int32),
        ("precision", C.c_int32), ("shift", C.c_int32), ("coefs", C.c_int32 * 32),
        ("coding_method", C.c_int32), ("partition_order", C.c_int32),
        ("rice", C.c_uint8 * 64), ("kind", C.c_uint8 * 64), ("bits", C.c_uint64),
    ]


class FrameInfo(C.Structure):
    _fields_ = [("channel_assignment", C.c_int32), ("channels", C.c_int32), ("frame_bytes", C.c_uint32),
                ("sub", SubframeInfo * 8)]


class Timings(C.Structure):
    _fields_ = [("total_ms", C.c_float), ("h2d_ms", C.c_float), ("d2h_ms", C.c_float), ("kernel_ms", C.c_float * 8),
                ("kernel_launches", C.c_uint32 * 8), ("launches", C.c_uint32)]


class WriterOptions(C.Structure):
    _fields_ = [("frame", Options), ("padding", C.c_int32), ("seektable_kind", C.c_uint32), ("seektable_n", C.c_uint32),
                ("launch_frames", C.c_uint32)]


class WriterStats(C.Structure):
    _fields_ = [("pcm_frames_written", C.c_uint64), ("frames_written", C.c_uint64), ("frame_bytes_written", C.c_uint64),
                ("min_frame_size", C.c_uint32), ("max_frame_size", C.c_uint32), ("launches", C.c_uint32),
                ("md5", C.c_uint8 * 16)]


class Streaminfo(C.Structure):
    _fields_ = [("min_block_size", C.c_uint16), ("max_block_size", C.c_uint16), ("min_frame_size", C.c_uint32),
                ("max_frame_size", C.c_uint32), ("sample_rate", C.c_uint32), ("channels", C.c_uint32),
                ("bits_per_sample", C.c_uint32), ("total_samples", C.c_uint64), ("md5", C.c_uint8 * 16),
                ("frames_start", C.c_uint64), ("n_seekpoints", C.c_uint32), ("reserved", C.c_uint32)]


class SeekPoint(C.Structure):
    _fields_ = [("sample_offset", C.c_uint64), ("byte_offset", C.c_uint64), ("frame_samples", C.c_uint32),
                ("placeholder", C.c_uint32)]


class FrameEntry(C.Structure):
    _fields_ = [("byte_offset", C.c_uint64), ("pcm_offset", C.c_uint64), ("byte_length", C.c_uint32), ("block_size", C.c_uint32)]


class FrameBuf(C.Structure):
    _fields_ = [("samples", C.POINTER(C.c_int32)), ("n_samples", C.c_size_t), ("sample_rate", C.c_uint32), ("channels", C.c_uint32),
                ("bits_per_sample", C.c_uint32), ("block_size", C.c_uint32)]


class Track(C.Structure):
    _fields_ = [("pcm", C.c_void_p), ("n_pcm_frames", C.c_uint64), ("sample_rate", C.c_uint32), ("bits_per_sample", C.c_uint32),
                ("channels", C.c_uint32), ("pcm_kind", C.c_int32)]


class File(C.Structure):
    _fields_ = [("data", C.c_void_p), ("capacity", C.c_size_t), ("len", C.c_size_t), ("status", C.c_int32), ("frames", C.c_uint32),
                ("md5", C.c_uint8 * 16)]


class Pcm(C.Structure):
    _fields_ = [("data", C.c_void_p), ("capacity", C.c_size_t), ("len", C.c_size_t), ("status", C.c_int32), ("verified", C.c_int32),
                ("info", Streaminfo)]


NEED_DATA = -10
NEED_SEEK = -11

EXPORTS = [
    "flacb200_writer_options_default", "flacb200_writer_options_fast", "flacb200_writer_options_best", "flacb200_writer_open",
    "flacb200_writer_close", "flacb200_total_from_bytes", "flacb200_total_from_samples", "flacb200_writer_header",
    "flacb200_writer_write_bytes", "flacb200_writer_write_samples", "flacb200_writer_write_channels", "flacb200_writer_drain",
    "flacb200_writer_flush", "flacb200_writer_finalize", "flacb200_writer_get_stats", "flacb200_read_streaminfo",
    "flacb200_reader_open", "flacb200_reader_close", "flacb200_reader_info", "flacb200_reader_seektable", "flacb200_reader_read",
    "flacb200_reader_seek", "flacb200_reader_verify", "flacb200_md5", "flacb200_md5_batch",
    "flacb200_reader_open_stream", "flacb200_reader_feed", "flacb200_reader_set_seekable", "flacb200_reader_wanted_offset", "flacb200_reader_set_window", "flacb200_reader_fill_buf",
    "flacb200_reader_consume", "flacb200_reader_fill_channels", "flacb200_reader_consume_channels", "flacb200_stream_write",
    "flacb200_stream_reader_open", "flacb200_stream_reader_close", "flacb200_stream_reader_feed", "flacb200_stream_reader_read",
    "flacb200_decode_last_frames", "flacb200_md5_many", "flacb200_encode_batch_bound", "flacb200_encode_batch", "flacb200_files_free",
    "flacb200_decode_batch", "flacb200_pcm_free", "flacb200_build_stream_header",
    "flacb200_options_default", "flacb200_options_fast", "flacb200_options_best", "flacb200_engine_create",
    "flacb200_engine_destroy", "flacb200_engine_set_stream", "flacb200_engine_set_chunk_frames", "flacb200_engine_set_keep_info", "flacb200_engine_set_option", "flacb200_encode",
    "flacb200_encode_bound", "flacb200_encode_last_info", "flacb200_decode", "flacb200_set_profiling",
    "flacb200_last_timings", "flacb200_debug_libm", "flacb200_synth_pcm", "flacb200_host_alloc", "flacb200_host_free",
    "flacb200_device_alloc", "flacb200_device_free", "flacb200_memcpy", "flacb200_synchronize", "flacb200_strerror",
    "flacb200_version",
]

_lib = None


class FlacB200Error(RuntimeError):
    def __init__(self, code: int, what: str = ""):
        self.code = code
        name = lib().flacb200_strerror(code).decode() if _lib is not None else str(code)
        super().__init__(f"{what}: error {code} ({name})")


def lib():
    """Load libflacb200.so (built in-tree by flac_codec_b200/build.py). Fails loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -m flac_codec_b200.build` (nvcc, sm_100a). "
                          "There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, u64p, u32p = C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)
    for n in ("flacb200_options_default", "flacb200_options_fast", "flacb200_options_best"):
        getattr(L, n).argtypes = [C.POINTER(Options)]
        getattr(L, n).restype = None
    L.flacb200_engine_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.flacb200_engine_destroy.argtypes = [vp]
    L.flacb200_engine_destroy.restype = None
    L.flacb200_engine_set_stream.argtypes = [vp, vp]
    L.flacb200_engine_set_chunk_frames.argtypes = [vp, C.c_uint32]
    L.flacb200_engine_set_keep_info.argtypes = [vp, C.c_int]
    L.flacb200_engine_set_option.argtypes = [vp, C.c_char_p, C.c_uint64]
    L.flacb200_encode.argtypes = [vp, C.POINTER(Options), C.POINTER(StreamParams), vp, C.c_size_t, C.c_int, C.c_int,
                                  C.c_uint64, C.POINTER(Segment), C.c_size_t, vp, C.c_size_t, C.c_int, u32p, C.c_size_t,
                                  u64p, u64p]
    L.flacb200_encode_bound.argtypes = [C.POINTER(Options), C.POINTER(StreamParams), C.POINTER(Segment), C.c_size_t]
    L.flacb200_encode_bound.restype = C.c_size_t
    L.flacb200_encode_last_info.argtypes = [vp, C.POINTER(FrameInfo), C.c_size_t, u64p]
    L.flacb200_decode.argtypes = [vp, C.POINTER(StreamParams), vp, C.c_size_t, C.c_int, C.POINTER(DecodeSegment),
                                  C.c_size_t, vp, C.c_size_t, C.c_int, C.c_int, C.c_uint64, u64p, u64p, u64p]
    L.flacb200_set_profiling.argtypes = [vp, C.c_int]
    L.flacb200_last_timings.argtypes = [vp, C.POINTER(Timings)]
    L.flacb200_synth_pcm.argtypes = [vp, vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32,
                                     C.c_uint64]
    L.flacb200_debug_libm.argtypes = [vp, C.c_int, vp, vp, C.c_size_t]
    L.flacb200_host_alloc.argtypes = [C.c_size_t]
    L.flacb200_host_alloc.restype = vp
    L.flacb200_host_free.argtypes = [vp]
    L.flacb200_host_free.restype = None
    L.flacb200_device_alloc.argtypes = [vp, C.c_size_t]
    L.flacb200_device_alloc.restype = vp
    L.flacb200_device_free.argtypes = [vp, vp]
    L.flacb200_device_free.restype = None
    L.flacb200_memcpy.argtypes = [vp, vp, vp, C.c_size_t, C.c_int]
    L.flacb200_synchronize.argtypes = [vp]
    L.flacb200_strerror.argtypes = [C.c_int]
    L.flacb200_strerror.restype = C.c_char_p
    L.flacb200_version.restype = C.c_char_p
    # ---- include/flacb200_stream.h ----
    u8pp, szp = C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_size_t)
    for n in ("flacb200_writer_options_default", "flacb200_writer_options_fast", "flacb200_writer_options_best"):
        getattr(L, n).argtypes = [C.POINTER(WriterOptions)]
        getattr(L, n).restype = None
    L.flacb200_writer_open.argtypes = [vp, C.POINTER(WriterOptions), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(vp)]
    L.flacb200_writer_close.argtypes = [vp]
    L.flacb200_writer_close.restype = None
    L.flacb200_total_from_bytes.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, u64p]
    L.flacb200_total_from_samples.argtypes = [C.c_uint64, C.c_uint32, u64p]
    L.flacb200_writer_header.argtypes = [vp, u8pp, szp]
    L.flacb200_writer_write_bytes.argtypes = [vp, vp, C.c_size_t, C.c_int]
    L.flacb200_writer_write_samples.argtypes = [vp, vp, C.c_size_t]
    L.flacb200_writer_write_channels.argtypes = [vp, C.POINTER(vp), C.c_uint32, C.c_size_t]
    L.flacb200_writer_drain.argtypes = [vp, u8pp, szp]
    L.flacb200_writer_flush.argtypes = [vp]
    L.flacb200_writer_finalize.argtypes = [vp]
    L.flacb200_writer_get_stats.argtypes = [vp, C.POINTER(WriterStats)]
    L.flacb200_read_streaminfo.argtypes = [vp, C.c_size_t, C.POINTER(Streaminfo)]
    L.flacb200_reader_open.argtypes = [vp, vp, C.c_size_t, C.POINTER(vp)]
    L.flacb200_reader_close.argtypes = [vp]
    L.flacb200_reader_close.restype = None
    L.flacb200_reader_info.argtypes = [vp, C.POINTER(Streaminfo)]
    L.flacb200_reader_seektable.argtypes = [vp, C.POINTER(SeekPoint), C.c_size_t, szp]
    L.flacb200_reader_read.argtypes = [vp, vp, C.c_size_t, C.c_int, szp]
    L.flacb200_reader_seek.argtypes = [vp, C.c_uint64]
    L.flacb200_reader_verify.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_uint8 * 16)]
    L.flacb200_reader_open_stream.argtypes = [vp, C.POINTER(vp)]
    L.flacb200_reader_feed.argtypes = [vp, vp, C.c_size_t, C.c_int]
    L.flacb200_reader_set_seekable.argtypes = [vp, C.c_int]
    L.flacb200_reader_wanted_offset.argtypes = [vp, u64p]
    L.flacb200_reader_set_window.argtypes = [vp, C.c_size_t, C.c_uint64]
    i32p = C.POINTER(C.c_int32)
    L.flacb200_reader_fill_buf.argtypes = [vp, C.POINTER(i32p), szp]
    L.flacb200_reader_consume.argtypes = [vp, C.c_size_t]
    L.flacb200_reader_fill_channels.argtypes = [vp, C.POINTER(C.POINTER(i32p)), szp]
    L.flacb200_reader_consume_channels.argtypes = [vp, C.c_size_t]
    L.flacb200_stream_write.argtypes = [vp, C.POINTER(Options), C.c_uint32, C.c_uint32, C.c_uint32, vp, C.c_size_t, C.c_uint64, vp, C.c_size_t, szp]
    L.flacb200_stream_reader_open.argtypes = [vp, C.POINTER(vp)]
    L.flacb200_stream_reader_close.argtypes = [vp]
    L.flacb200_stream_reader_close.restype = None
    L.flacb200_stream_reader_feed.argtypes = [vp, vp, C.c_size_t, C.c_int]
    L.flacb200_stream_reader_read.argtypes = [vp, C.POINTER(FrameBuf)]
    L.flacb200_decode_last_frames.argtypes = [vp, C.POINTER(FrameEntry), C.c_size_t, u64p]
    L.flacb200_md5_many.argtypes = [C.POINTER(vp), szp, C.c_size_t, vp, C.c_uint]
    L.flacb200_md5_many.restype = None
    L.flacb200_encode_batch_bound.argtypes = [C.POINTER(Track), C.POINTER(WriterOptions)]
    L.flacb200_encode_batch_bound.restype = C.c_size_t
    L.flacb200_encode_batch.argtypes = [C.POINTER(Track), C.c_size_t, C.POINTER(WriterOptions), C.POINTER(C.c_int), C.c_int, C.POINTER(File)]
    L.flacb200_files_free.argtypes = [C.POINTER(File), C.c_size_t]
    L.flacb200_files_free.restype = None
    L.flacb200_decode_batch.argtypes = [C.POINTER(vp), szp, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(Pcm)]
    L.flacb200_pcm_free.argtypes = [C.POINTER(Pcm), C.c_size_t]
    L.flacb200_pcm_free.restype = None
    L.flacb200_md5.argtypes = [vp, C.c_size_t, C.POINTER(C.c_uint8 * 16)]
    L.flacb200_md5.restype = None
    L.flacb200_md5_batch.argtypes = [vp, vp, C.c_size_t, C.c_int, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(Segment), C.c_size_t,
                                     C.POINTER(C.c_uint8)]
    L.flacb200_md5_batch.restype = C.c_int
    _lib = L
    return L


def check(code: int, what: str):
    if code != 0:
        raise FlacB200Error(code, what)
