"""BASELINE.json's configurations other than the bench workload (C4), each at its STATED size: run through the C ABI /
the stream facades on the GPU and checked against the CPU oracle.  Test infrastructure: tests/test_gpu_configs.py
asserts on the results, bench.py adds them to its JSON line as `configs` (the oracle is the checker here, never the
thing measured).

  C1  wav2flac-equivalent: 60 s 44.1 kHz/16-bit stereo, Options::default(), whole .flac file (metadata, seek table,
      MD5, frames) == the oracle's file
  C2  flac2wav-equivalent: that file decoded through FlacByteReader == the input PCM == the oracle's decoder
  C3  60 s 96 kHz/24-bit 8-channel, Options::best(): all 1407 frames byte-identical
  C5  192 kHz/32-bit streams at max LPC order 32, blocks 4096 and 16384, 2 and 8 channels (encode byte-identical, decode
      bit-exact, both decoders), plus the decoder over the reference's own .flac fixtures (the IETF testbench the config
      names is not in the reference tree)

Every leg returns {"name", "msamples_per_s", "identical", "frames", ...}; times are wall clock around the call with HOST
buffers (best of `reps`), single-channel samples per second.
"""
from __future__ import annotations

import hashlib
import io
import os
import time

import numpy as np

from flacb200_testutil import ref_file, synth_pcm

C5_SHAPES = ((2, 4096, 30), (2, 16384, 30), (8, 4096, 10), (8, 16384, 10))   # channels, block size, seconds


class _Pinned:
    """Pinned host buffers for the batch-ABI legs (pageable memory would time the driver's staging copies)."""

    def __init__(self):
        from flac_codec_b200 import _abi

        self.L, self.ptrs = _abi.lib(), []

    def buf(self, nbytes, src=None):
        import ctypes

        p = self.L.flacb200_host_alloc(nbytes)
        if not p:
            raise MemoryError("flacb200_host_alloc")
        self.ptrs.append(p)
        a = np.ctypeslib.as_array((ctypes.c_uint8 * nbytes).from_address(p))
        if src is not None:
            a[:] = src
        return p, a

    def free(self):
        for p in self.ptrs:
            self.L.flacb200_host_free(p)
        self.ptrs = []


def _encode_pinned(eng, opt, rate, bps, ch, raw, n, reps):
    """eng.encode of one stream with pinned PCM in and pinned frames out; returns (seconds, frame bytes, sizes, total)."""
    from flac_codec_b200 import _abi

    pin = _Pinned()
    try:
        hp, _ = pin.buf(raw.nbytes, raw)
        cap = raw.nbytes + raw.nbytes // 8 + (1 << 20)
        ho, hout = pin.buf(cap)

        def enc():
            return eng.encode(opt, rate, bps, ch, hp, raw.nbytes, _abi.PCM_BYTES_LE, [(0, n, 0)], out=ho, out_capacity=cap)

        enc()
        t, (_, sizes, total) = _best(enc, reps)
        eng.set_profiling(True)   # one more call with the engine's own stage clock: what the wall time is made of
        try:
            enc()
            tm = eng.timings()
            _encode_pinned.engine_ms = {"h2d": round(float(tm.h2d_ms), 3), "planes": round(float(tm.kernel_ms[0]), 3),
                                        "lpc": round(float(tm.kernel_ms[1]), 3), "analysis": round(float(tm.kernel_ms[2]), 3),
                                        "decide_scan": round(float(tm.kernel_ms[3]), 3), "pack": round(float(tm.kernel_ms[4]), 3),
                                        "d2h": round(float(tm.d2h_ms), 3)}
        finally:
            eng.set_profiling(False)
        return t, hout[:total].tobytes(), sizes, total
    finally:
        pin.free()


def _best(fn, reps):
    best, out = 1e30, None
    for _ in range(max(reps, 1)):
        t0 = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t0)
    return best, out


def c1_c2(eng, fo, reps=2, seconds=60):
    from flac_codec_b200 import Options, stream

    rate, bps, ch = 44100, 16, 2
    x = synth_pcm(0, ch, rate * seconds, rate, bps).reshape(-1)
    raw = fo.samples_to_bytes(x, 2)
    cores = os.cpu_count() or 1
    ref, ref_sizes = fo.encode_stream(fo.options("default"), rate, bps, ch, x, total_known=True, nthreads=cores)

    def enc():
        sink = io.BytesIO()
        w = stream.FlacByteWriter(sink, Options.default(), rate, bps, ch, len(raw), engine=eng, launch_frames=1024)
        w.write(raw)
        w.finalize()
        w.close()
        return sink.getvalue()

    enc()
    t, flac = _best(enc, reps)
    si = fo.read_streaminfo(flac)
    legs = [{"name": f"C1 wav2flac {seconds} s 44.1k/16/2 default, whole file via FlacByteWriter", "msamples_per_s": x.size / t / 1e6,
             "ms": t * 1e3, "identical": flac == ref, "frames": int(len(ref_sizes)), "file_bytes": len(flac),
             "md5_ok": bytes(si.md5) == hashlib.md5(raw).digest()}]

    out = np.zeros(len(raw), dtype=np.uint8)   # caller-owned, already touched: the leg times the reader, not first-touch page faults

    def dec():
        r = stream.FlacByteReader(flac, engine=eng)
        n = r.readinto(out)
        r.close()
        return n

    dec()
    t, got = _best(dec, reps)
    want, _ = fo.decode_stream(ref)
    legs.append({"name": "C2 flac2wav of the C1 file via FlacByteReader", "msamples_per_s": x.size / t / 1e6, "ms": t * 1e3,
                 "identical": got == len(raw) and out.tobytes() == raw and np.array_equal(want, x), "frames": int(len(ref_sizes))})
    return legs


def c3(eng, fo, reps=2, seconds=60):
    from flac_codec_b200 import Options, _abi

    rate, bps, ch = 96000, 24, 8
    n = rate * seconds
    x = synth_pcm(3, ch, n, rate, bps).reshape(-1)
    raw = np.frombuffer(fo.samples_to_bytes(x, 3), dtype=np.uint8).copy()
    cores = os.cpu_count() or 1
    ref, ref_sizes = fo.encode_frames_only(fo.options("best"), rate, bps, ch, x, nthreads=cores)

    t, data, sizes, total = _encode_pinned(eng, Options.best(), rate, bps, ch, raw, n, reps)
    ident = data == ref
    same = int((np.asarray(sizes) == np.asarray(ref_sizes)).sum()) if len(sizes) == len(ref_sizes) else 0
    return [{"name": f"C3 {seconds} s 96k/24/8ch best, all frames", "msamples_per_s": x.size / t / 1e6, "ms": t * 1e3, "identical": ident,
             "frames": int(len(ref_sizes)), "equal_frame_sizes": same, "size_delta": (total - len(ref)) / len(ref),
             "engine_ms": _encode_pinned.engine_ms}]


def c5(eng, fo, reps=1, shapes=C5_SHAPES):
    from flac_codec_b200 import Options, _abi, stream

    rate, bps = 192000, 32
    cores = os.cpu_count() or 1
    legs = []
    for ch, block, seconds in shapes:
        n = rate * seconds
        x = synth_pcm(5, ch, n, rate, bps).reshape(-1)
        raw = np.frombuffer(fo.samples_to_bytes(x, 4), dtype=np.uint8).copy()
        opt5 = fo.options("best", max_lpc_order=32, block_size=block)
        flac5, sizes5 = fo.encode_stream(opt5, rate, bps, ch, x, total_known=True, nthreads=cores)
        ref, _ = fo.encode_frames_only(opt5, rate, bps, ch, x, nthreads=cores)
        o5 = Options.best().max_lpc_order(32).block_size(block)

        t, data, _, total = _encode_pinned(eng, o5, rate, bps, ch, raw, n, reps)
        legs.append({"name": f"C5 encode {seconds} s 192k/32/{ch}ch LPC<=32 block {block}", "msamples_per_s": x.size / t / 1e6,
                     "ms": t * 1e3, "identical": data == ref, "frames": int(len(sizes5)), "engine_ms": _encode_pinned.engine_ms})
        for legacy, label in ((0, "k_parse+k_restore"), (64, "k_decode")):
            eng.set_option("legacy", legacy)
            y = np.zeros(x.size + 16, dtype=np.int32)   # caller-owned and touched
            try:
                def dec():
                    r = stream.FlacSampleReader(flac5, engine=eng)
                    n = r.readinto(y)
                    r.close()
                    return n

                dec()
                t, got = _best(dec, reps)
            finally:
                eng.set_option("legacy", 0)
            legs.append({"name": f"C5 decode {seconds} s 192k/32/{ch}ch order 32 block {block} ({label})",
                         "msamples_per_s": x.size / t / 1e6, "ms": t * 1e3, "identical": bool(got == x.size and np.array_equal(y[: x.size], x)),
                         "frames": int(len(sizes5))})
    # the reference's own .flac fixtures (what stands in for the decoder testbench here)
    for name in ("sine.flac", "all-frames.flac", "cuesheet.flac", "seektable.flac"):
        flac = ref_file(name)
        si = fo.read_streaminfo(flac)

        nbytes = si.total_samples * si.channels * ((si.bps + 7) // 8)
        out = np.zeros(nbytes + 16, dtype=np.uint8)

        def dec():
            r = stream.FlacByteReader(flac, engine=eng)
            n = r.readinto(out)
            r.close()
            return n

        dec()
        t, got = _best(dec, reps)
        legs.append({"name": f"C5 decode fixture {name}", "msamples_per_s": si.total_samples * si.channels / t / 1e6, "ms": t * 1e3,
                     "identical": got == nbytes and hashlib.md5(out[:nbytes].tobytes()).digest() == bytes(si.md5), "frames": None})
    return legs


def all_legs(eng, fo, reps=2):
    return c1_c2(eng, fo, reps) + c3(eng, fo, reps) + c5(eng, fo, 1)
