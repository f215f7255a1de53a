"""Python mirror of flac-codec's writer/reader facades over the stream-level C ABI (include/flacb200_stream.h).

Same type names, argument meaning and error behaviour as the reference (src/encode.rs:103-1290,
src/decode.rs:103-1309) so that its tests can be replayed against the GPU engine.  The byte sink/source is any
binary file object (`write`/`seek`/`tell` for writers -- the reference's `W: Write + Seek`; `read` for readers).
All work happens in libflacb200.so; nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _abi
from ._abi import FlacB200Error, check
from .engine import Engine, Options

_ENGINES: dict = {}


def default_engine(device: int = 0) -> Engine:
    """One shared engine per device for the facades (an engine is one CUDA stream + scratch)."""
    if device not in _ENGINES:
        _ENGINES[device] = Engine(device)
    return _ENGINES[device]


def _writer_options(opt: Options, launch_frames: int = 0) -> _abi.WriterOptions:
    wo = _abi.WriterOptions()
    wo.frame = opt.c
    wo.padding = -1 if not opt.padding else int(opt.padding)
    if opt.seektable is None:
        wo.seektable_kind, wo.seektable_n = 0, 0
    else:
        wo.seektable_kind = 1 if opt.seektable[0] == "seconds" else 2
        wo.seektable_n = int(opt.seektable[1])
        if wo.seektable_n == 0:   # NonZero::new(0) -> None (:1572, :1580)
            wo.seektable_kind = 0
    wo.launch_frames = launch_frames
    return wo


class _Writer:
    """Encoder (src/encode.rs:1860-2110) behind flacb200_writer."""

    def __init__(self, writer, options: Options, sample_rate: int, bits_per_sample: int, channels: int,
                 total_pcm_frames: int, *, engine: Optional[Engine] = None, launch_frames: int = 0):
        self._L = _abi.lib()
        self._sink = writer
        self._h = C.c_void_p()
        self._engine = engine if engine is not None else default_engine()
        wo = _writer_options(options, launch_frames)
        check(self._L.flacb200_writer_open(self._engine._h, C.byref(wo), sample_rate, bits_per_sample, channels,
                                           total_pcm_frames, C.byref(self._h)), "Encoder::new")
        self._start = writer.tell()          # writer.stream_position() (:1941)
        self._sink.write(self._header())     # write_blocks (:1953)
        self._finalized = False
        self.channels, self.bits_per_sample, self.sample_rate = channels, bits_per_sample, sample_rate

    def _header(self) -> bytes:
        p, n = C.POINTER(C.c_uint8)(), C.c_size_t(0)
        check(self._L.flacb200_writer_header(self._h, C.byref(p), C.byref(n)), "writer_header")
        return C.string_at(p, n.value)

    def _drain(self):
        p, n = C.POINTER(C.c_uint8)(), C.c_size_t(0)
        check(self._L.flacb200_writer_drain(self._h, C.byref(p), C.byref(n)), "writer_drain")
        if n.value:
            self._sink.write(C.string_at(p, n.value))

    def _after(self, rc: int, what: str):
        if rc == 0:
            self._drain()
        check(rc, what)

    def flush(self):
        self._after(self._L.flacb200_writer_flush(self._h), "flush")
        if hasattr(self._sink, "flush"):
            self._sink.flush()

    def finalize(self):
        """finalize_inner (:2024): last block, sample-count check, MD5, seek table, metadata rewrite."""
        if self._finalized:
            return
        self._finalized = True
        rc = self._L.flacb200_writer_finalize(self._h)
        self._after(rc, "finalize")
        end = self._sink.tell()
        self._sink.seek(self._start)
        self._sink.write(self._header())
        self._sink.seek(end)

    def stats(self) -> _abi.WriterStats:
        s = _abi.WriterStats()
        check(self._L.flacb200_writer_get_stats(self._h, C.byref(s)), "stats")
        return s

    def close(self):
        if self._h:
            self._L.flacb200_writer_close(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        try:
            if exc[0] is None:
                self.finalize()
        finally:
            self.close()

    def __del__(self):   # Drop finalises and swallows errors (:399-405)
        try:
            if self._h and not self._finalized:
                self.finalize()
        except Exception:
            pass
        try:
            self.close()
        except Exception:
            pass


def _total(fn, *args) -> int:
    out = C.c_uint64(0)
    check(fn(*args, C.byref(out)), "total")
    return out.value


class FlacByteWriter(_Writer):
    """FlacByteWriter<W, E> (src/encode.rs:103-405): interleaved PCM bytes in, a .flac stream out."""

    def __init__(self, writer, options: Options, sample_rate: int, bits_per_sample: int, channels: int,
                 total_bytes: Optional[int] = None, *, endian: str = "little", **kw):
        L = _abi.lib()
        if not 1 <= bits_per_sample <= 32:
            raise FlacB200Error(33, "FlacByteWriter::new")
        total = 0 if total_bytes is None else _total(L.flacb200_total_from_bytes, total_bytes, bits_per_sample, channels)
        self._big = {"little": 0, "big": 1}[endian]
        super().__init__(writer, options, sample_rate, bits_per_sample, channels, total, **kw)

    @classmethod
    def new_cdda(cls, writer, options: Options, total_bytes: Optional[int] = None, **kw):
        return cls(writer, options, 44100, 16, 2, total_bytes, **kw)

    def write(self, buf) -> int:   # io::Write::write always consumes the whole slice (:347)
        a = np.frombuffer(buf, dtype=np.uint8)
        self._after(self._L.flacb200_writer_write_bytes(self._h, C.c_void_p(a.ctypes.data if a.size else 0), a.size, self._big),
                    "FlacByteWriter::write")
        return a.size


class FlacSampleWriter(_Writer):
    """FlacSampleWriter<W> (src/encode.rs:431-628): interleaved i32 samples in."""

    def __init__(self, writer, options: Options, sample_rate: int, bits_per_sample: int, channels: int,
                 total_samples: Optional[int] = None, **kw):
        L = _abi.lib()
        if not 1 <= bits_per_sample <= 32:
            raise FlacB200Error(33, "FlacSampleWriter::new")
        total = 0 if total_samples is None else _total(L.flacb200_total_from_samples, total_samples, channels)
        super().__init__(writer, options, sample_rate, bits_per_sample, channels, total, **kw)

    @classmethod
    def new_cdda(cls, writer, options: Options, total_samples: Optional[int] = None, **kw):
        return cls(writer, options, 44100, 16, 2, total_samples, **kw)

    def write(self, samples):
        a = np.ascontiguousarray(samples, dtype=np.int32).reshape(-1)
        self._after(self._L.flacb200_writer_write_samples(self._h, C.c_void_p(a.ctypes.data if a.size else 0), a.size),
                    "FlacSampleWriter::write")


class FlacChannelWriter(_Writer):
    """FlacChannelWriter<W> (src/encode.rs:713-893): one slice per channel; total_samples counts per channel."""

    def __init__(self, writer, options: Options, sample_rate: int, bits_per_sample: int, channels: int,
                 total_samples: Optional[int] = None, **kw):
        if not 1 <= bits_per_sample <= 32:
            raise FlacB200Error(33, "FlacChannelWriter::new")
        if total_samples == 0:
            raise FlacB200Error(63, "FlacChannelWriter::new")   # InvalidTotalSamples
        super().__init__(writer, options, sample_rate, bits_per_sample, channels, total_samples or 0, **kw)

    def write(self, channels: Sequence):
        chans = [np.ascontiguousarray(c, dtype=np.int32).reshape(-1) for c in channels]
        if len(chans) != self.channels:
            raise FlacB200Error(64, "FlacChannelWriter::write")     # ChannelCountMismatch (:845)
        n = chans[0].size if chans else 0
        if any(c.size != n for c in chans):
            raise FlacB200Error(65, "FlacChannelWriter::write")     # ChannelLengthMismatch (:851)
        ptrs = (C.c_void_p * len(chans))(*[c.ctypes.data for c in chans])
        self._after(self._L.flacb200_writer_write_channels(self._h, ptrs, len(chans), n), "FlacChannelWriter::write")


class FlacStreamWriter:
    """FlacStreamWriter<W> (src/encode.rs:1063-1290): subset frames only, no metadata, parameters per call."""

    def __init__(self, writer, options: Options, *, engine: Optional[Engine] = None):
        self._sink, self._opt = writer, options
        self._engine = engine if engine is not None else default_engine()
        self._frame_number = 0

    def write(self, sample_rate: int, channels: int, bits_per_sample: int, samples):
        a = np.ascontiguousarray(samples, dtype=np.int32).reshape(-1)
        if channels == 0 or a.size % channels:
            raise FlacB200Error(61, "FlacStreamWriter::write")   # SamplesNotDivisibleByChannels (:1103)
        n = a.size // channels
        if n == 0:
            return
        if n > 65535:
            raise FlacB200Error(24, "FlacStreamWriter::write")   # InvalidBlockSize: one frame per call (:1118)
        from .engine import Options as _O   # a copy with the block size of this call

        opt = _O("default")
        C.memmove(C.byref(opt.c), C.byref(self._opt.c), C.sizeof(opt.c))
        opt.c.block_size = max(n, 1)
        data, sizes, total = self._engine.encode(opt, sample_rate, bits_per_sample, channels, a, a.nbytes,
                                                 _abi.PCM_I32_INTERLEAVED, [(0, n, self._frame_number)], subset=True)
        self._frame_number += 1
        self._sink.write(data.tobytes())

    def write_cdda(self, samples):
        self.write(44100, 2, 16, samples)


# ---------------------------------------------------------------------------------------------------------------
class _Reader:
    def __init__(self, reader, *, engine: Optional[Engine] = None):
        self._L = _abi.lib()
        data = reader if isinstance(reader, (bytes, bytearray, memoryview)) else reader.read()
        self._image = np.frombuffer(bytes(data), dtype=np.uint8)   # kept alive: the handle borrows it
        self._engine = engine if engine is not None else default_engine()
        self._h = C.c_void_p()
        check(self._L.flacb200_reader_open(self._engine._h, C.c_void_p(self._image.ctypes.data), self._image.size,
                                           C.byref(self._h)), "FlacReader::new")
        self._si = _abi.Streaminfo()
        check(self._L.flacb200_reader_info(self._h, C.byref(self._si)), "reader_info")

    # Metadata trait (src/metadata/mod.rs:48-105)
    def channel_count(self) -> int:
        return self._si.channels

    def sample_rate(self) -> int:
        return self._si.sample_rate

    def bits_per_sample(self) -> int:
        return self._si.bits_per_sample

    def total_samples(self) -> Optional[int]:
        return self._si.total_samples or None

    def md5(self) -> Optional[bytes]:
        m = bytes(self._si.md5)
        return m if any(m) else None

    def decoded_len(self) -> Optional[int]:
        t = self.total_samples()
        return None if t is None else t * self._si.channels * ((self._si.bits_per_sample + 7) // 8)

    def seektable(self):
        n = C.c_size_t(0)
        check(self._L.flacb200_reader_seektable(self._h, None, 0, C.byref(n)), "seektable")
        pts = (_abi.SeekPoint * max(n.value, 1))()
        check(self._L.flacb200_reader_seektable(self._h, pts, n.value, C.byref(n)), "seektable")
        return [(p.sample_offset, p.byte_offset, p.frame_samples, bool(p.placeholder)) for p in pts[: n.value]]

    def seek(self, sample: int):   # Decoder::seek (src/decode.rs:1452): inter-channel sample index
        check(self._L.flacb200_reader_seek(self._h, sample), "seek")

    def verify(self):
        res, md5 = C.c_int(0), (C.c_uint8 * 16)()
        check(self._L.flacb200_reader_verify(self._h, C.byref(res), C.byref(md5)), "verify")
        return ("MD5Match", "MD5Mismatch", "NoMD5")[res.value], bytes(md5)

    def close(self):
        if self._h:
            self._L.flacb200_reader_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FlacByteReader(_Reader):
    """FlacByteReader<R, E> (src/decode.rs:103-371)."""

    def __init__(self, reader, *, endian: str = "little", **kw):
        super().__init__(reader, **kw)
        self._kind = _abi.PCM_BYTES_LE if endian == "little" else _abi.PCM_BYTES_BE

    def read(self, size: int = -1) -> bytes:
        chunks = []
        want = size if size >= 0 else 1 << 62
        known = self.decoded_len()
        while want > 0:
            # one call for the whole stream when its length is known (a fresh 16 MB buffer per call costs more than the decode)
            cap = min(want, max(known, 1) if known is not None and size < 0 and not chunks else 1 << 24)
            buf = np.empty(cap, dtype=np.uint8)
            n = C.c_size_t(0)
            check(self._L.flacb200_reader_read(self._h, C.c_void_p(buf.ctypes.data), cap, self._kind, C.byref(n)), "read")
            if n.value == 0:
                break
            chunks.append(buf[: n.value].tobytes())
            want -= n.value
        return b"".join(chunks)


class FlacSampleReader(_Reader):
    """FlacSampleReader<R> (src/decode.rs:384-620)."""

    def read(self, n_samples: int) -> np.ndarray:
        buf = np.empty(max(n_samples, 1), dtype=np.int32)
        n = C.c_size_t(0)
        check(self._L.flacb200_reader_read(self._h, C.c_void_p(buf.ctypes.data), n_samples, _abi.PCM_I32_INTERLEAVED,
                                           C.byref(n)), "read")
        return buf[: n.value]

    def read_to_end(self) -> np.ndarray:
        out = []
        while True:
            a = self.read(1 << 22)
            if a.size == 0:
                break
            out.append(a.copy())
        return np.concatenate(out) if out else np.zeros(0, dtype=np.int32)


class FlacStreamReader:
    """FlacStreamReader<R> (src/decode.rs:1158-1268): subset frames without metadata; all frames of the image are
    decoded in one batch and handed out one FrameBuf at a time."""

    def __init__(self, reader, sample_rate: int, channels: int, bits_per_sample: int, *, engine: Optional[Engine] = None):
        data = reader if isinstance(reader, (bytes, bytearray, memoryview)) else reader.read()
        self._image = np.frombuffer(bytes(data), dtype=np.uint8)
        self._engine = engine if engine is not None else default_engine()
        self.sample_rate, self.channels, self.bits_per_sample = sample_rate, channels, bits_per_sample

    def read_all(self, max_pcm_frames: int) -> np.ndarray:
        out = np.zeros(max_pcm_frames * self.channels, dtype=np.int32)
        nf, ns = self._engine.decode(self.sample_rate, self.bits_per_sample, self.channels, 0, self._image, self._image.size,
                                     [(0, self._image.size, 0, 0)], out, out.nbytes, _abi.PCM_I32_INTERLEAVED, subset=True)
        return out[: ns * self.channels]


def verify(reader, **kw):
    """flac_codec::decode::verify_reader (src/decode.rs:1291)."""
    r = FlacByteReader(reader, **kw)
    try:
        return r.verify()[0]
    finally:
        r.close()
