#!/usr/bin/env python
"""End-to-end decode (host frames -> host PCM) of the bench workload for several batch sizes of decode_batched."""
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
    L = _abi.lib()
    h_pcm = L.flacb200_host_alloc(nbytes); h_out = L.flacb200_host_alloc(cap)
    segs = [(t * n, n, 0) for t in range(ntr)]
    _, sizes, total = eng.encode(Options.best(), rate, bps, ch, d_pcm, nbytes, _abi.PCM_BYTES_LE, segs, pcm_location=_abi.DEVICE, out=h_out,
                                 out_capacity=cap, out_location=_abi.HOST, want_sizes=True)
    eng.device_free(d_pcm)
    per = (n + 4095) // 4096
    offs = np.concatenate([[0], np.cumsum(sizes.astype(np.int64))])
    dsegs = [(int(offs[t * per]), int(offs[(t + 1) * per] - offs[t * per]), t * n, n) for t in range(ntr)]
    res = {}
    for mb in (0, 96, 192, 384, 768, 1536):
        if mb == 0:
            eng.set_option("no_batch", 1)
        else:
            eng.set_option("no_batch", 0)
            eng.set_option("batch_bytes", mb << 20)
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            nf, ns = eng.decode(rate, bps, ch, 4096, h_out, total, dsegs, h_pcm, nbytes, _abi.PCM_BYTES_LE, frames_location=_abi.HOST,
                                pcm_location=_abi.HOST)
            best = min(best, time.perf_counter() - t0)
        res["one call" if mb == 0 else f"{mb} MB batches"] = round(best * 1e3, 1)
    print(json.dumps(res))

if __name__ == "__main__":
    main()
