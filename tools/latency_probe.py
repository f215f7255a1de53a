"""Where the time of ONE stream goes (the single-stream latency question of BASELINE's C1 / C3 / C5 configs).

For each shape: wall clock of flacb200_encode with pinned host buffers and with device-resident buffers, plus the engine's own
stage clock (flacb200_last_timings with profiling on: upload, planes, lpc, analysis, decide+scan, pack, download).  One JSON
line per shape.  Not a parity tool: tests/test_gpu_configs.py holds the identical-to-the-oracle checks of the same shapes.
"""
from __future__ import annotations

import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from flac_codec_b200 import Engine, Options, _abi  # noqa: E402

SHAPES = (   # name, rate, bps, channels, seconds, options
    ("C1 60 s 44.1k/16/2 default", 44100, 16, 2, 60, lambda: Options.default()),
    ("C1 60 s 44.1k/16/2 best", 44100, 16, 2, 60, lambda: Options.best()),
    ("C3 60 s 96k/24/8 best", 96000, 24, 8, 60, lambda: Options.best()),
    ("C3-like 60 s 96k/24/2 best", 96000, 24, 2, 60, lambda: Options.best()),
    ("C5 30 s 192k/32/2 order 32 block 4096", 192000, 32, 2, 30, lambda: Options.best().max_lpc_order(32).block_size(4096)),
    ("C5 10 s 192k/32/8 order 32 block 16384", 192000, 32, 8, 10, lambda: Options.best().max_lpc_order(32).block_size(16384)),
)
STAGES = ("planes", "lpc", "analysis", "decide_scan", "pack")


def best(fn, reps=5):
    out, t = None, 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        t = min(t, time.perf_counter() - t0)
    return t, out


def main():
    eng = Engine(0)
    L = _abi.lib()
    for name, rate, bps, ch, secs, mk in SHAPES:
        opt = mk()
        n = rate * secs
        bytes_ps = (bps + 7) // 8
        nbytes = n * ch * bytes_ps
        d_pcm = eng.device_alloc(nbytes)
        eng.synth_pcm(d_pcm, 0, 1, n, ch, rate, bps)
        hp = L.flacb200_host_alloc(nbytes)
        eng.memcpy(hp, d_pcm, nbytes, 2)
        cap = nbytes + nbytes // 8 + (1 << 20)
        ho = L.flacb200_host_alloc(cap)
        d_out = eng.device_alloc(cap)

        def host():
            return eng.encode(opt, rate, bps, ch, hp, nbytes, _abi.PCM_BYTES_LE, [(0, n, 0)], pcm_location=_abi.HOST, out=ho, out_capacity=cap,
                              out_location=_abi.HOST)

        def resident():
            return eng.encode(opt, rate, bps, ch, d_pcm, nbytes, _abi.PCM_BYTES_LE, [(0, n, 0)], pcm_location=_abi.DEVICE, out=d_out,
                              out_capacity=cap, out_location=_abi.DEVICE)

        host(), resident()
        t_host, (_, sizes, total) = best(host)
        t_res, _ = best(resident)
        eng.set_profiling(True)
        resident()
        tm = eng.timings()
        host()
        tmh = eng.timings()
        eng.set_profiling(False)
        line = {"shape": name, "frames": int(len(sizes)), "samples": n * ch, "pcm_mb": nbytes / 1e6, "flac_mb": total / 1e6,
                "host_ms": t_host * 1e3, "resident_ms": t_res * 1e3, "launches": int(tm.launches),
                "resident_stages_ms": {s: round(float(tm.kernel_ms[i]), 3) for i, s in enumerate(STAGES)},
                "resident_kernels_ms": round(float(tm.total_ms), 3),
                "host_stages_ms": {s: round(float(tmh.kernel_ms[i]), 3) for i, s in enumerate(STAGES)},
                "host_h2d_ms": round(float(tmh.h2d_ms), 3), "host_d2h_ms": round(float(tmh.d2h_ms), 3),
                "host_msamples_per_s": n * ch / t_host / 1e6, "resident_msamples_per_s": n * ch / t_res / 1e6}
        # decode of the same frames: host frames -> host PCM (pinned), then device-resident
        hq = L.flacb200_host_alloc(nbytes + 64)
        d_back = eng.device_alloc(nbytes + 64)
        eng.memcpy(d_out, ho, total, 1)
        bsz = opt.c.block_size

        def dec_host():
            return eng.decode(rate, bps, ch, bsz, ho, total, [(0, total, 0, n)], hq, nbytes, _abi.PCM_BYTES_LE)

        def dec_res():
            return eng.decode(rate, bps, ch, bsz, d_out, total, [(0, total, 0, n)], d_back, nbytes, _abi.PCM_BYTES_LE,
                              frames_location=_abi.DEVICE, pcm_location=_abi.DEVICE)

        dec_host(), dec_res()
        td_host, (nf, ns) = best(dec_host)
        td_res, _ = best(dec_res)
        same = bytes((ctypes.c_uint8 * nbytes).from_address(hq)) == bytes((ctypes.c_uint8 * nbytes).from_address(hp))
        eng.set_profiling(True)
        dec_res()
        td = eng.timings()
        eng.set_profiling(False)
        line["decode"] = {"host_ms": td_host * 1e3, "resident_ms": td_res * 1e3, "frames": int(nf), "bit_exact": bool(same and ns == n),
                          "launches": int(td.launches), "resident_stages_ms": [round(float(v), 3) for v in td.kernel_ms[:6]],
                          "host_msamples_per_s": n * ch / td_host / 1e6, "resident_msamples_per_s": n * ch / td_res / 1e6}
        L.flacb200_host_free(hq)
        eng.device_free(d_back)
        print(json.dumps(line), flush=True)
        eng.device_free(d_pcm)
        eng.device_free(d_out)
        L.flacb200_host_free(hp)
        L.flacb200_host_free(ho)


if __name__ == "__main__":
    main()
