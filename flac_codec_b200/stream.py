"""Python mirror of flac-codec's writer/reader facades over the stream-level C ABI (include/flacb200_stream.h).

Same type names, argument meaning and error behaviour as the reference (src/encode.rs:103-1290,
src/decode.rs:103-1309) so that its tests can be replayed against the GPU engine.  The byte sink/source is any
binary file object (`write`/`seek`/`tell` for writers -- the reference's `W: Write + Seek`; `read` for readers).
All work happens in libflacb200.so; nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
import io
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
    """FlacStreamWriter<W> (src/encode.rs:1063-1290): subset frames only, no metadata, parameters per call
    (flacb200_stream_write)."""

    def __init__(self, writer, options: Options, *, engine: Optional[Engine] = None):
        self._sink, self._opt = writer, options
        self._engine = engine if engine is not None else default_engine()
        self._frame_number = 0

    def write(self, sample_rate: int, channels: int, bits_per_sample: int, samples):
        a = np.ascontiguousarray(samples, dtype=np.int32).reshape(-1)
        cap = a.size * 5 + 1024
        out = np.empty(cap, dtype=np.uint8)
        n = C.c_size_t(0)
        check(_abi.lib().flacb200_stream_write(self._engine._h, C.byref(self._opt.c), sample_rate, channels, bits_per_sample,
                                               C.c_void_p(a.ctypes.data if a.size else 0), a.size, self._frame_number,
                                               C.c_void_p(out.ctypes.data), cap, C.byref(n)), "FlacStreamWriter::write")
        if n.value:
            self._frame_number += 1
            self._sink.write(out[: n.value].tobytes())

    def write_cdda(self, samples):
        self.write(44100, 2, 16, samples)


# ---------------------------------------------------------------------------------------------------------------
class _Reader:
    """Decoder (src/decode.rs:1311-1491) behind flacb200_reader.

    `reader` is a bytes-like file image or a file object with read() [+ seek()].  A seekable source (new_seekable / open)
    is held as an image; with `streaming=True` (the reference's plain `new(R: Read)`) the bytes are fed to the handle in
    chunks of `chunk` bytes as the decoder asks for them, and the reader cannot seek."""

    def __init__(self, reader, *, engine: Optional[Engine] = None, streaming: bool = False, chunk: int = 1 << 20,
                 window_bytes: int = 0, window_pcm_frames: int = 0, seekable: bool = False):
        self._L = _abi.lib()
        self._engine = engine if engine is not None else default_engine()
        self._h = C.c_void_p()
        self._src, self._chunk, self._eof = None, chunk, False
        if streaming:
            self._src = io.BytesIO(bytes(reader)) if isinstance(reader, (bytes, bytearray, memoryview)) else reader
            check(self._L.flacb200_reader_open_stream(self._engine._h, C.byref(self._h)), "FlacReader::new")
            if seekable:   # new_seekable over a source that stays on disk: seeks reposition the source, nothing is held
                check(self._L.flacb200_reader_set_seekable(self._h, 1), "set_seekable")
        else:
            data = reader if isinstance(reader, (bytes, bytearray, memoryview)) else reader.read()
            self._image = np.frombuffer(bytes(data), dtype=np.uint8)   # kept alive: the handle borrows it
            check(self._L.flacb200_reader_open(self._engine._h, C.c_void_p(self._image.ctypes.data), self._image.size,
                                               C.byref(self._h)), "FlacReader::new")
        if window_bytes or window_pcm_frames:
            check(self._L.flacb200_reader_set_window(self._h, window_bytes, window_pcm_frames), "set_window")
        self._si = _abi.Streaminfo()
        self._call(lambda: self._L.flacb200_reader_info(self._h, C.byref(self._si)), "BlockList::read")

    def _call(self, fn, what: str):
        """Runs one handle call; in streaming mode FLACB200_NEED_DATA is answered by feeding the next chunk of the source."""
        while True:
            rc = fn()
            if rc == _abi.NEED_SEEK:
                off = C.c_uint64(0)
                check(self._L.flacb200_reader_wanted_offset(self._h, C.byref(off)), "wanted_offset")
                self._src.seek(off.value)
                self._eof = False
                continue
            if rc != _abi.NEED_DATA:
                check(rc, what)
                return
            if self._src is None or self._eof:
                check(1, what)   # Io: the source ended inside a frame
            b = self._src.read(self._chunk)
            a = np.frombuffer(b, dtype=np.uint8)
            self._eof = a.size == 0
            check(self._L.flacb200_reader_feed(self._h, C.c_void_p(a.ctypes.data if a.size else 0), a.size, int(self._eof)), "feed")

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

    def seek(self, sample: int):
        """FlacSampleReader::seek / FlacChannelReader::seek (src/decode.rs:823, :1021): inter-channel sample index."""
        self._call(lambda: self._L.flacb200_reader_seek(self._h, sample), "seek")

    def verify(self):
        res, md5 = C.c_int(0), (C.c_uint8 * 16)()
        self._call(lambda: self._L.flacb200_reader_verify(self._h, C.byref(res), C.byref(md5)), "verify")
        return ("MD5Match", "MD5Mismatch", "NoMD5")[res.value], bytes(md5)

    def _read(self, buf: np.ndarray, capacity: int, kind: int) -> int:
        n = C.c_size_t(0)
        self._call(lambda: self._L.flacb200_reader_read(self._h, C.c_void_p(buf.ctypes.data), capacity, kind, C.byref(n)), "read")
        return n.value

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
    """FlacByteReader<R, E> (src/decode.rs:103-371): io::Read + io::BufRead + io::Seek over the decoded PCM bytes."""

    def __init__(self, reader, *, endian: str = "little", **kw):
        super().__init__(reader, **kw)
        self._kind = _abi.PCM_BYTES_LE if endian == "little" else _abi.PCM_BYTES_BE
        self._pos = 0   # byte position, for io::Seek

    def read(self, size: int = -1) -> bytes:
        chunks = []
        want = size if size >= 0 else 1 << 62
        known = self.decoded_len()
        while want > 0:
            # one call for the whole stream when its length is known (a fresh 16 MB buffer per call costs more than the decode)
            cap = min(want, max(known, 1) if known is not None and size < 0 and not chunks else 1 << 24)
            buf = np.empty(cap, dtype=np.uint8)
            n = self._read(buf, cap, self._kind)
            if n == 0:
                break
            chunks.append(buf[:n].tobytes())
            want -= n
        out = b"".join(chunks)
        self._pos += len(out)
        return out

    def readinto(self, buf) -> int:
        """io::Read::read into a caller-owned buffer (numpy uint8 array / writable buffer): no intermediate copies."""
        a = buf if isinstance(buf, np.ndarray) else np.frombuffer(buf, dtype=np.uint8)
        got = 0
        while got < a.size:
            n = self._read(a[got:], a.size - got, self._kind)
            if n == 0:
                break
            got += n
        self._pos += got
        return got

    def seek_bytes(self, offset: int, whence: int = 0) -> int:
        """io::Seek (src/decode.rs:715-820): byte positions of the decoded stream.  The decoder seeks to the PCM frame that
        holds the byte, then bytes are consumed up to the position itself (the next read continues inside that PCM frame)."""
        bpf = ((self._si.bits_per_sample + 7) // 8) * self._si.channels
        if whence == 0:
            want = offset
        elif whence == 1:
            want = self._pos + offset
        else:
            if not self._si.total_samples:
                raise OSError("total samples not known")
            if offset > 0:
                raise OSError("cannot seek beyond end of file")
            # (the reference takes total_samples -- PCM frames, not bytes -- as the end position, :766-768; kept as it is)
            want = self._si.total_samples + offset
        if want < 0:
            raise OSError("cannot seek below byte 0")
        self.seek(want // bpf)
        self._pos = (want // bpf) * bpf
        skip = want - self._pos
        if skip:
            got = self.read(skip)   # (advances self._pos)
            if len(got) < skip:
                raise EOFError("stream exhausted before sample reached")
        return want


class FlacSampleReader(_Reader):
    """FlacSampleReader<R> (src/decode.rs:384-620)."""

    def read(self, n_samples: int) -> np.ndarray:
        buf = np.empty(max(n_samples, 1), dtype=np.int32)
        return buf[: self._read(buf, n_samples, _abi.PCM_I32_INTERLEAVED)]

    def readinto(self, buf: np.ndarray) -> int:
        """Fills a caller-owned int32 array with interleaved samples; returns how many were read."""
        got = 0
        while got < buf.size:
            n = self._read(buf[got:], buf.size - got, _abi.PCM_I32_INTERLEAVED)
            if n == 0:
                break
            got += n
        return got

    def read_to_end(self) -> np.ndarray:
        total = self._si.total_samples * self._si.channels
        if total:   # announced length: one buffer, filled in place (no per-chunk copies); an understated total falls through
            buf = np.empty(total, dtype=np.int32)
            got = 0
            while got < total:
                n = self._read(buf[got:], total - got, _abi.PCM_I32_INTERLEAVED)
                if n == 0:
                    break
                got += n
            if got < total:
                return buf[:got]
            rest = self.read(1 << 16)
            if rest.size == 0:
                return buf
            out = [buf, rest.copy()]
        else:
            out = []
        while True:
            a = self.read(1 << 22)
            if a.size == 0:
                break
            out.append(a.copy())
        return np.concatenate(out) if out else np.zeros(0, dtype=np.int32)

    def fill_buf(self) -> np.ndarray:
        """The unconsumed interleaved samples of the current frame (:466); empty at the end of the stream."""
        p, n = C.POINTER(C.c_int32)(), C.c_size_t(0)
        self._call(lambda: self._L.flacb200_reader_fill_buf(self._h, C.byref(p), C.byref(n)), "fill_buf")
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy() if n.value else np.zeros(0, dtype=np.int32)

    def consume(self, amt: int):
        check(self._L.flacb200_reader_consume(self._h, amt), "consume")

    def __iter__(self):
        """FlacSampleIterator (:667-712): one sample at a time."""
        while True:
            buf = self.fill_buf()
            if buf.size == 0:
                return
            self.consume(buf.size)
            yield from (int(v) for v in buf)


class FlacChannelReader(_Reader):
    """FlacChannelReader<R> (src/decode.rs:880-1065): one slice per channel."""

    def fill_buf(self):
        pp, n = C.POINTER(C.POINTER(C.c_int32))(), C.c_size_t(0)
        self._call(lambda: self._L.flacb200_reader_fill_channels(self._h, C.byref(pp), C.byref(n)), "fill_buf")
        return [np.ctypeslib.as_array(pp[c], shape=(n.value,)).copy() if n.value else np.zeros(0, dtype=np.int32)
                for c in range(self._si.channels)]

    def consume(self, amt: int):
        check(self._L.flacb200_reader_consume_channels(self._h, amt), "consume")


class FlacStreamReader:
    """FlacStreamReader<R> (src/decode.rs:1149-1268): subset frames without metadata.  read() returns the next frame as
    (samples, sample_rate, channels, bits_per_sample) -- FrameBuf -- with the parameters of that frame's own header."""

    def __init__(self, reader, *, engine: Optional[Engine] = None, chunk: int = 1 << 20):
        self._L = _abi.lib()
        self._engine = engine if engine is not None else default_engine()
        self._src = io.BytesIO(bytes(reader)) if isinstance(reader, (bytes, bytearray, memoryview)) else reader
        self._chunk, self._eof = chunk, False
        self._h = C.c_void_p()
        check(self._L.flacb200_stream_reader_open(self._engine._h, C.byref(self._h)), "FlacStreamReader::new")

    def read(self):
        fb = _abi.FrameBuf()
        while True:
            rc = self._L.flacb200_stream_reader_read(self._h, C.byref(fb))
            if rc != _abi.NEED_DATA:
                check(rc, "FlacStreamReader::read")
                break
            if self._eof:
                check(1, "FlacStreamReader::read")
            b = self._src.read(self._chunk)
            a = np.frombuffer(b, dtype=np.uint8)
            self._eof = a.size == 0
            check(self._L.flacb200_stream_reader_feed(self._h, C.c_void_p(a.ctypes.data if a.size else 0), a.size, int(self._eof)), "feed")
        samples = np.ctypeslib.as_array(fb.samples, shape=(fb.n_samples,)).copy()
        return samples, fb.sample_rate, fb.channels, fb.bits_per_sample

    def close(self):
        if self._h:
            self._L.flacb200_stream_reader_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def verify(reader, **kw):
    """flac_codec::decode::verify_reader (src/decode.rs:1291)."""
    r = FlacByteReader(reader, **kw)
    try:
        return r.verify()[0]
    finally:
        r.close()
