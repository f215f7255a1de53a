#!/usr/bin/env python
"""Where a single-stream decode spends its time: per-kernel CUDA-event times of flacb200_decode for the C2 stream
(60 s 44.1k/16/2) and the C5 streams (10 s 192k/32/2, LPC order 32), beside the wall clock of the call and of the
FlacSampleReader facade.  Diagnostic only (executes oracle/ to make the input streams)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from flac_codec_b200 import Engine, _abi, stream  # noqa: E402
from flacb200_testutil import synth_pcm  # noqa: E402
from oracle import oracle as fo  # noqa: E402

DN = ["k_find+k_scan", "k_parse", "k_restore || k_crc16f+k_chain_fast", None, "k_emit"]


def run(eng, name, rate, bps, ch, secs, opt, block):
    x = synth_pcm(5, ch, rate * secs, rate, bps).reshape(-1)
    frames, sizes = fo.encode_frames_only(opt, rate, bps, ch, x, nthreads=os.cpu_count())
    fr = np.frombuffer(frames, dtype=np.uint8).copy()
    out = np.zeros(x.size, dtype=np.int32)
    segs = [(0, fr.nbytes, 0, rate * secs)]
    eng.set_profiling(True)
    best = 1e9
    for _ in range(4):
        t0 = time.perf_counter()
        nf, ns = eng.decode(rate, bps, ch, block, fr, fr.nbytes, segs, out, out.nbytes, _abi.PCM_I32_INTERLEAVED)
        best = min(best, time.perf_counter() - t0)
    tm = eng.timings()
    assert np.array_equal(out, x)
    print(json.dumps({"stream": name, "frames": int(nf), "call_ms": best * 1e3,
                      "kernel_ms": {DN[k]: round(tm.kernel_ms[k], 3) for k in range(5) if DN[k]}, "launches": tm.launches, "h2d_ms": tm.h2d_ms, "d2h_ms": tm.d2h_ms,
                      "msamples_per_s": x.size / best / 1e6}), flush=True)


def main():
    eng = Engine(0)
    run(eng, "C2 60s 44.1k/16/2 default", 44100, 16, 2, 60, fo.options("default"), 4096)
    run(eng, "C5 10s 192k/32/2 order32 b4096", 192000, 32, 2, 10, fo.options("best", max_lpc_order=32, block_size=4096), 4096)
    run(eng, "C5 10s 192k/32/2 order32 b16384", 192000, 32, 2, 10, fo.options("best", max_lpc_order=32, block_size=16384), 16384)
    run(eng, "C3 10s 96k/24/8 best", 96000, 24, 8, 10, fo.options("best"), 4096)


if __name__ == "__main__":
    main()
