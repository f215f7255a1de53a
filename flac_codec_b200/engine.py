"""Thin object wrapper over the C ABI: one Engine per GPU (mirrors flacb200_engine)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _abi
from ._abi import (DEVICE, HOST, PCM_BYTES_BE, PCM_BYTES_LE, PCM_I32_INTERLEAVED, PCM_I32_PLANAR, DecodeSegment,
                   FrameInfo, Segment, StreamParams, Timings, check)


class Options:
    """Mirror of flac_codec::encode::Options (src/encode.rs:1363-1672): the frame-engine fields plus the
    container fields the host-side writers use.  Builder methods keep the reference's names and checks."""

    def __init__(self, preset: str = "default"):
        self._o = _abi.Options()
        getattr(_abi.lib(), f"flacb200_options_{preset}")(C.byref(self._o))
        self.padding: Optional[int] = 4096          # Options::default() inserts a 4096-byte PADDING (:1392)
        self.seektable: Optional[tuple] = ("seconds", 10)  # SeekTableInterval::default() (:1329)

    @classmethod
    def default(cls):
        return cls("default")

    @classmethod
    def fast(cls):
        return cls("fast")

    @classmethod
    def best(cls):
        return cls("best")

    def block_size(self, n: int):   # :1418  OptionsError::InvalidBlockSize
        if n < 16 or n > 65535:
            raise ValueError("block size must be >= 16")
        self._o.block_size = n
        return self

    def max_lpc_order(self, n: Optional[int]):   # :1430  OptionsError::InvalidLpcOrder
        if n is not None and not (0 < n <= 32):
            raise ValueError("maximum LPC order must be <= 32")
        self._o.max_lpc_order = n or 0
        return self

    def max_partition_order(self, n: int):   # :1447  OptionsError::InvalidMaxPartitions
        if not (0 <= n <= 15):
            raise ValueError("max partition order must be <= 15")
        self._o.max_partition_order = n
        return self

    def mid_side(self, on: bool):
        self._o.mid_side = int(on)
        return self

    def fast_channel_correlation(self, fast: bool):
        self._o.exhaustive_channel_correlation = int(not fast)
        return self

    def window(self, kind: str, p: float = 0.5):
        self._o.window_kind = {"rectangle": 0, "hann": 1, "tukey": 2}[kind]
        self._o.tukey_p = p
        return self

    def no_padding(self):
        self.padding = None
        return self

    def with_padding(self, size: int):
        self.padding = size
        return self

    def no_seektable(self):
        self.seektable = None
        return self

    def seektable_seconds(self, s: int):
        self.seektable = ("seconds", s)
        return self

    def seektable_frames(self, n: int):
        self.seektable = ("frames", n)
        return self

    @property
    def c(self) -> _abi.Options:
        return self._o


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return C.c_void_p(a.ctypes.data)


class Engine:
    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        check(_abi.lib().flacb200_engine_create(device, C.byref(self._h)), "flacb200_engine_create")
        self.device = device

    def close(self):
        if self._h:
            _abi.lib().flacb200_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- configuration --
    def set_stream(self, cuda_stream: int):
        check(_abi.lib().flacb200_engine_set_stream(self._h, C.c_void_p(cuda_stream)), "set_stream")

    def set_chunk_frames(self, n: int):
        check(_abi.lib().flacb200_engine_set_chunk_frames(self._h, n), "set_chunk_frames")

    def set_keep_info(self, on: bool):
        check(_abi.lib().flacb200_engine_set_keep_info(self._h, int(on)), "set_keep_info")

    def set_option(self, key: str, value: int):
        """Runtime knobs ("legacy", "batch_bytes", "no_batch", "debug"); the environment is only read at creation."""
        check(_abi.lib().flacb200_engine_set_option(self._h, key.encode(), int(value)), f"set_option({key})")

    def set_profiling(self, on: bool):
        check(_abi.lib().flacb200_set_profiling(self._h, int(on)), "set_profiling")

    def timings(self) -> Timings:
        t = Timings()
        check(_abi.lib().flacb200_last_timings(self._h, C.byref(t)), "last_timings")
        return t

    def debug_libm(self, fn: int, x: np.ndarray) -> np.ndarray:
        """Device-side glibc_log (fn 0) / glibc_log2 (fn 1) of a float64 array (parity tooling)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty_like(x)
        check(_abi.lib().flacb200_debug_libm(self._h, fn, _ptr(x), _ptr(out), x.size), "debug_libm")
        return out

    def synchronize(self):
        check(_abi.lib().flacb200_synchronize(self._h), "synchronize")

    # -- memory helpers (device pointers are plain ints) --
    def device_alloc(self, nbytes: int) -> int:
        p = _abi.lib().flacb200_device_alloc(self._h, nbytes)
        if not p:
            raise MemoryError(f"flacb200_device_alloc({nbytes})")
        return p

    def device_free(self, p: int):
        _abi.lib().flacb200_device_free(self._h, C.c_void_p(p))

    def memcpy(self, dst, src, nbytes: int, kind: int):
        check(_abi.lib().flacb200_memcpy(self._h, _ptr(dst), _ptr(src), nbytes, kind), "memcpy")

    def synth_pcm(self, dptr: int, first_track: int, n_tracks: int, n_pcm_frames: int, channels: int, rate: int,
                  bps: int, seed: int = 20261017):
        check(_abi.lib().flacb200_synth_pcm(self._h, C.c_void_p(dptr), first_track, n_tracks, n_pcm_frames, channels, rate,
                                            bps, seed), "synth_pcm")

    # -- encode --
    def encode(self, opt: Options, rate: int, bps: int, channels: int, pcm, pcm_bytes: int, pcm_kind: int,
               segments: Sequence[tuple], *, pcm_location: int = HOST, out=None, out_capacity: int = 0,
               out_location: int = HOST, subset: bool = False, planar_stride: int = 0, want_sizes: bool = True):
        """Batch encode (flacb200_encode). segments: (pcm_offset, n_pcm_frames, first_frame_number).
        Returns (out_bytes_or_None, frame_sizes, total_bytes)."""
        L = _abi.lib()
        segs = (Segment * len(segments))(*[Segment(*s) for s in segments])
        params = StreamParams(rate, bps, channels, int(subset), 0, 0)
        nframes = sum((s[1] + opt.c.block_size - 1) // opt.c.block_size for s in segments)
        own = None
        if out is None and out_location == HOST:
            cap = L.flacb200_encode_bound(C.byref(opt.c), C.byref(params), segs, len(segments))
            own = np.empty(cap, dtype=np.uint8)
            out, out_capacity = own, cap
        sizes = np.zeros(max(nframes, 1), dtype=np.uint32) if want_sizes else None
        nf, total = C.c_uint64(0), C.c_uint64(0)
        rc = L.flacb200_encode(self._h, C.byref(opt.c), C.byref(params), _ptr(pcm), pcm_bytes, pcm_kind, pcm_location,
                               planar_stride, segs, len(segments), _ptr(out), out_capacity, out_location,
                               sizes.ctypes.data_as(C.POINTER(C.c_uint32)) if sizes is not None else None,
                               sizes.size if sizes is not None else 0, C.byref(nf), C.byref(total))
        check(rc, "flacb200_encode")
        data = own[: total.value] if own is not None else None
        return data, (sizes[: nf.value] if sizes is not None else None), total.value

    def last_info(self):
        L = _abi.lib()
        n = C.c_uint64(0)
        check(L.flacb200_encode_last_info(self._h, None, 0, C.byref(n)), "last_info")
        infos = (FrameInfo * max(n.value, 1))()
        check(L.flacb200_encode_last_info(self._h, infos, n.value, C.byref(n)), "last_info")
        return infos, n.value

    # -- MD5 of many streams (update_md5, src/encode.rs:1292; verify, src/decode.rs:1291) --
    def md5(self, bps: int, channels: int, pcm, pcm_bytes: int, pcm_kind: int, segments: Sequence[tuple], *,
            pcm_location: int = HOST, planar_stride: int = 0):
        """segments: (pcm_offset, n_pcm_frames).  Returns one 16-byte digest per segment."""
        segs = (Segment * len(segments))(*[Segment(s[0], s[1], 0) for s in segments])
        out = np.zeros(16 * max(len(segments), 1), dtype=np.uint8)
        check(_abi.lib().flacb200_md5_batch(self._h, _ptr(pcm), pcm_bytes, pcm_kind, pcm_location, planar_stride, channels, bps, segs,
                                            len(segments), out.ctypes.data_as(C.POINTER(C.c_uint8))), "flacb200_md5_batch")
        return [out[16 * i: 16 * i + 16].tobytes() for i in range(len(segments))]

    # -- decode --
    def decode(self, rate: int, bps: int, channels: int, max_block_size: int, frames, frames_bytes: int,
               segments: Sequence[tuple], pcm_out, pcm_out_bytes: int, pcm_kind: int, *, frames_location: int = HOST,
               pcm_location: int = HOST, subset: bool = False, planar_stride: int = 0):
        """Batch decode (flacb200_decode). segments: (byte_offset, byte_length, pcm_offset, n_pcm_frames).
        Returns (n_frames, n_pcm_frames)."""
        L = _abi.lib()
        segs = (DecodeSegment * len(segments))(*[DecodeSegment(*s) for s in segments])
        params = StreamParams(rate, bps, channels, int(subset), max_block_size, 0)
        nf, ns, bad = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        rc = L.flacb200_decode(self._h, C.byref(params), _ptr(frames), frames_bytes, frames_location, segs, len(segments),
                               _ptr(pcm_out), pcm_out_bytes, pcm_kind, pcm_location, planar_stride, C.byref(nf), C.byref(ns),
                               C.byref(bad))
        if rc != 0:
            err = _abi.FlacB200Error(rc, f"flacb200_decode (frame {bad.value})")
            err.bad_frame = bad.value
            raise err
        return nf.value, ns.value
