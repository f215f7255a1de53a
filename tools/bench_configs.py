#!/usr/bin/env python
"""Side benchmark over the BASELINE.json configs that bench.py does not time (C1, C2, C3, C5) -- one JSON line each.

  C1  wav2flac-equivalent: 60 s 44.1 kHz/16-bit stereo, Options::default(), whole .flac file through FlacByteWriter
  C2  flac2wav-equivalent: decode of that file through FlacByteReader, bit-exact
  C3  60 s 96 kHz/24-bit 8-channel, Options::best()
  C5  decode of 192 kHz/32-bit streams at LPC order 32 (oracle-encoded; generic kernels), bit-exact
plus batched variants (many such streams per call) that show what the same kernels do when the GPU is full.
Times are wall clock around the C-ABI / facade call (host buffers in, host buffers out), best of 3.
The oracle (CPU restatement) is timed beside each on one thread and on all cores.
"""
from __future__ import annotations

import io
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from flac_codec_b200 import Engine, Options, _abi, stream  # noqa: E402
from flacb200_testutil import synth_pcm  # noqa: E402
from oracle import oracle as fo  # noqa: E402


def best_of(fn, n=3):
    best, out = 1e30, None
    for _ in range(n):
        t0 = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t0)
    return best, out


def pinned(eng, nbytes):
    p = _abi.lib().flacb200_host_alloc(nbytes)
    return p, np.ctypeslib.as_array((__import__("ctypes").c_uint8 * nbytes).from_address(p))


def main():
    eng = Engine(0)
    cores = os.cpu_count() or 1
    out = []

    def emit(d):
        print(json.dumps(d), flush=True)
        out.append(d)

    # ---------------- C1 / C2 ----------------
    rate, bps, ch, secs = 44100, 16, 2, 60
    x = synth_pcm(0, ch, rate * secs, rate, bps).reshape(-1)
    raw = fo.samples_to_bytes(x, 2)
    nsamp = x.size

    def c1():
        sink = io.BytesIO()
        w = stream.FlacByteWriter(sink, Options.default(), rate, bps, ch, len(raw), engine=eng, launch_frames=1024)
        w.write(raw)
        w.finalize()
        w.close()
        return sink.getvalue()

    c1()
    t_gpu, flac = best_of(c1)
    t_cpu1, (ref, _) = best_of(lambda: fo.encode_stream(fo.options("default"), rate, bps, ch, x, total_known=True, nthreads=1), 1)
    t_cpuN, _ = best_of(lambda: fo.encode_stream(fo.options("default"), rate, bps, ch, x, total_known=True, nthreads=cores), 2)
    emit({"config": "C1 wav2flac 60 s 44.1k/16/2 default (one stream, whole file through FlacByteWriter)",
          "gpu_msamples_per_s": nsamp / t_gpu / 1e6, "gpu_ms": t_gpu * 1e3, "cpu_1thread_msamples_per_s": nsamp / t_cpu1 / 1e6,
          f"cpu_{cores}threads_msamples_per_s": nsamp / t_cpuN / 1e6, "file_identical_to_oracle": flac == ref,
          "size_delta": (len(flac) - len(ref)) / len(ref), "flac_bytes": len(flac)})

    def c2():
        r = stream.FlacByteReader(flac, engine=eng)
        b = r.read()
        r.close()
        return b

    c2()
    t_gpu, pcm = best_of(c2)
    t_cpu1, _ = best_of(lambda: fo.decode_stream(ref), 1)
    emit({"config": "C2 flac2wav of the C1 stream (FlacByteReader)", "gpu_msamples_per_s": nsamp / t_gpu / 1e6, "gpu_ms": t_gpu * 1e3,
          "cpu_1thread_msamples_per_s": nsamp / t_cpu1 / 1e6, "bit_exact": pcm == raw})

    # batched C1: 256 such tracks per call through the batch C ABI, pinned host buffers
    ntr, n = 256, rate * secs
    nbytes = ntr * n * ch * 2
    d_pcm = eng.device_alloc(nbytes)
    eng.synth_pcm(d_pcm, 0, ntr, n, ch, rate, bps)
    hp, h_pcm = pinned(eng, nbytes)
    eng.memcpy(hp, d_pcm, nbytes, 2)
    cap = nbytes + nbytes // 8 + (1 << 20)
    ho, h_out = pinned(eng, cap)
    segs = [(t * n, n, 0) for t in range(ntr)]
    eng.set_keep_info(False)

    def c1b():
        return eng.encode(Options.default(), rate, bps, ch, hp, nbytes, _abi.PCM_BYTES_LE, segs, pcm_location=_abi.HOST, out=ho,
                          out_capacity=cap, out_location=_abi.HOST)

    c1b()
    t_gpu, (_, sizes, total) = best_of(c1b)
    per = (n + 4095) // 4096
    ref0, ref_sizes0 = fo.encode_frames_only(fo.options("default"), rate, bps, ch, fo.bytes_to_samples(h_pcm[: n * ch * 2].tobytes(), 2), nthreads=cores)
    emit({"config": f"C1 batched: {ntr} tracks x 60 s 44.1k/16/2 default per call (host PCM -> host frames)",
          "gpu_msamples_per_s": ntr * n * ch / t_gpu / 1e6, "gpu_ms": t_gpu * 1e3, "ratio": total / nbytes,
          "first_track_identical_to_oracle": bytes(h_out[: int(sizes[:per].sum())]) == ref0})

    offs = np.concatenate([[0], np.cumsum(sizes.astype(np.int64))])
    dsegs = [(int(offs[t * per]), int(offs[(t + 1) * per] - offs[t * per]), t * n, n) for t in range(ntr)]
    hb, h_back = pinned(eng, nbytes)

    def c2b():
        return eng.decode(rate, bps, ch, 4096, ho, total, dsegs, hb, nbytes, _abi.PCM_BYTES_LE, frames_location=_abi.HOST,
                          pcm_location=_abi.HOST)

    c2b()
    t_gpu, _ = best_of(c2b)
    emit({"config": f"C2 batched: decode of those {ntr} streams per call (host frames -> host PCM)",
          "gpu_msamples_per_s": ntr * n * ch / t_gpu / 1e6, "gpu_ms": t_gpu * 1e3, "bit_exact": bool(np.array_equal(h_back, h_pcm))})
    eng.device_free(d_pcm)
    for p in (hp, ho, hb):
        _abi.lib().flacb200_host_free(p)

    # ---------------- C3 ----------------
    rate, bps, ch, secs = 96000, 24, 8, 60
    n = rate * secs
    nbytes = n * ch * 3
    d_pcm = eng.device_alloc(nbytes)
    eng.synth_pcm(d_pcm, 0, 1, n, ch, rate, bps)
    hp, h_pcm = pinned(eng, nbytes)
    eng.memcpy(hp, d_pcm, nbytes, 2)
    cap = nbytes + nbytes // 8 + (1 << 20)
    ho, h_out = pinned(eng, cap)

    def c3():
        return eng.encode(Options.best(), rate, bps, ch, hp, nbytes, _abi.PCM_BYTES_LE, [(0, n, 0)], pcm_location=_abi.HOST, out=ho,
                          out_capacity=cap, out_location=_abi.HOST)

    c3()
    t_gpu, (_, sizes, total) = best_of(c3)
    x3 = fo.bytes_to_samples(h_pcm[: 10 * rate * ch * 3].tobytes(), 3)   # first 10 s on the CPU
    t_cpu1, (ref3, rs3) = best_of(lambda: fo.encode_frames_only(fo.options("best"), rate, bps, ch, x3, nthreads=1), 1)
    t_cpuN, _ = best_of(lambda: fo.encode_frames_only(fo.options("best"), rate, bps, ch, x3, nthreads=cores), 2)
    nf3 = len(rs3) - 1   # whole blocks of the 10 s prefix
    same = bytes(h_out[: int(sizes[:nf3].sum())]) == ref3[: int(rs3[:nf3].sum())]
    emit({"config": "C3 60 s 96k/24/8ch best (one stream, host PCM -> host frames)", "gpu_msamples_per_s": n * ch / t_gpu / 1e6,
          "gpu_ms": t_gpu * 1e3, "cpu_1thread_msamples_per_s": x3.size / t_cpu1 / 1e6,
          f"cpu_{cores}threads_msamples_per_s": x3.size / t_cpuN / 1e6, "ratio": total / nbytes, "first_10s_identical_to_oracle": same})
    eng.device_free(d_pcm)
    for p in (hp, ho):
        _abi.lib().flacb200_host_free(p)

    # ---------------- C5 ----------------
    for block in (4096, 16384):
        rate, bps, ch, secs = 192000, 32, 2, 10
        x5 = synth_pcm(5, ch, rate * secs, rate, bps).reshape(-1)
        opt5 = fo.options("best", max_lpc_order=32, block_size=block)
        flac5, _ = fo.encode_stream(opt5, rate, bps, ch, x5, total_known=True, nthreads=cores)

        def c5():
            r = stream.FlacSampleReader(flac5, engine=eng)
            y = r.read_to_end()
            r.close()
            return y

        c5()
        t_gpu, y = best_of(c5)
        t_cpu1, _ = best_of(lambda: fo.decode_stream(flac5), 1)
        emit({"config": f"C5 decode 10 s 192k/32/2, LPC order 32, block {block} (one stream, FlacSampleReader)",
              "gpu_msamples_per_s": x5.size / t_gpu / 1e6, "gpu_ms": t_gpu * 1e3, "cpu_1thread_msamples_per_s": x5.size / t_cpu1 / 1e6,
              "bit_exact": bool(np.array_equal(y, x5))})
        # and the GPU encoder at that shape (generic kernels: 32-bit samples, order 32)
        raw5 = np.frombuffer(fo.samples_to_bytes(x5, 4), dtype=np.uint8)
        o5 = Options.best().max_lpc_order(32).block_size(block)
        eng.set_keep_info(False)
        t_enc, (data5, sizes5, total5) = best_of(lambda: eng.encode(o5, rate, bps, ch, raw5, raw5.nbytes, _abi.PCM_BYTES_LE, [(0, rate * secs, 0)]), 2)
        ref5, _ = fo.encode_frames_only(opt5, rate, bps, ch, x5, nthreads=cores)
        emit({"config": f"C5 encode of the same PCM (generic kernels), block {block}", "gpu_msamples_per_s": x5.size / t_enc / 1e6,
              "gpu_ms": t_enc * 1e3, "identical_to_oracle": data5.tobytes() == ref5})
    with open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
