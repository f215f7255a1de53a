#!/usr/bin/env python
"""Throughput of the GPU many-stream MD5 (flacb200_md5_batch) on the bench shape: 128 tracks x 180 s 48 kHz/24-bit
stereo resident in HBM, next to the host MD5 of stream.cpp (one thread) -- one JSON line."""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flac_codec_b200 import Engine, _abi  # noqa: E402


def main():
    eng = Engine(0)
    rate, bps, ch, ntr, n = 48000, 24, 2, int(os.environ.get("TRACKS", "128")), 48000 * 180
    nbytes = ntr * n * ch * 3
    d = eng.device_alloc(nbytes)
    eng.synth_pcm(d, 0, ntr, n, ch, rate, bps)
    segs = [(t * n, n) for t in range(ntr)]
    eng.md5(bps, ch, d, nbytes, _abi.PCM_BYTES_LE, segs, pcm_location=_abi.DEVICE)
    t0 = time.perf_counter()
    got = eng.md5(bps, ch, d, nbytes, _abi.PCM_BYTES_LE, segs, pcm_location=_abi.DEVICE)
    dt = time.perf_counter() - t0
    kernel_ms = eng.timings().total_ms
    per = n * ch * 3
    host = np.zeros(per, dtype=np.uint8)
    eng.memcpy(host, d, per, 2)
    t0 = time.perf_counter()
    want = hashlib.md5(host.tobytes()).digest()
    t_host = time.perf_counter() - t0
    print(json.dumps({"tracks": ntr, "bytes": nbytes, "gpu_ms": dt * 1e3, "kernel_ms": kernel_ms, "gpu_gbs": nbytes / dt / 1e9,
                      "per_stream_mbs": per / (kernel_ms * 1e-3) / 1e6, "msamples_per_s": ntr * n * ch / dt / 1e6,
                      "first_digest_ok": got[0] == want, "hashlib_one_thread_mbs": per / t_host / 1e6}))
    eng.device_free(d)


if __name__ == "__main__":
    main()
