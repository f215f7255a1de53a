#!/usr/bin/env python
"""Where the end-to-end encode step goes: the bench workload with host/device combinations of input and output."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flac_codec_b200 import Engine, Options, _abi

def main():
    eng = Engine(0); eng.set_keep_info(False)
    rate, bps, ch, ntr, n = 48000, 24, 2, 128, 48000 * 180
    nbytes = ntr * n * ch * 3
    d_pcm = eng.device_alloc(nbytes); eng.synth_pcm(d_pcm, 0, ntr, n, ch, rate, bps)
    cap = nbytes + nbytes // 8 + (1 << 20)
    d_out = eng.device_alloc(cap)
    L = _abi.lib()
    h_pcm = L.flacb200_host_alloc(nbytes); h_out = L.flacb200_host_alloc(cap)
    eng.memcpy(h_pcm, d_pcm, nbytes, 2)
    segs = [(t * n, n, 0) for t in range(ntr)]
    opt = Options.best()
    res = {}
    for chunk in (0, 4096, 16384):
        eng.set_chunk_frames(chunk)
        for name, (pi, pl, po, ol) in {"host->host": (h_pcm, _abi.HOST, h_out, _abi.HOST), "host->device": (h_pcm, _abi.HOST, d_out, _abi.DEVICE),
                                       "device->host": (d_pcm, _abi.DEVICE, h_out, _abi.HOST), "device->device": (d_pcm, _abi.DEVICE, d_out, _abi.DEVICE)}.items():
            best = 1e9
            for _ in range(3):
                t0 = time.perf_counter()
                eng.encode(opt, rate, bps, ch, pi, nbytes, _abi.PCM_BYTES_LE, segs, pcm_location=pl, out=po, out_capacity=cap, out_location=ol, want_sizes=True)
                best = min(best, time.perf_counter() - t0)
            res[f"chunk={chunk} {name}"] = round(best * 1e3, 2)
    print(json.dumps(res, indent=1))

if __name__ == "__main__":
    main()
