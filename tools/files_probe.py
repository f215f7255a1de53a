#!/usr/bin/env python
"""Where the whole-file batch step goes: flacb200_md5_many alone at several thread counts, flacb200_encode_batch as it is."""
import ctypes as C, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flac_codec_b200 import Engine, Options, _abi
from flac_codec_b200.batch import encode_files

def main():
    eng = Engine(0)
    rate, bps, ch, ntr, n = 48000, 24, 2, int(os.environ.get("TRACKS", "128")), 48000 * 180
    tb = n * ch * 3
    nbytes = ntr * tb
    L = _abi.lib()
    d = eng.device_alloc(nbytes); eng.synth_pcm(d, 0, ntr, n, ch, rate, bps)
    hp = L.flacb200_host_alloc(nbytes); eng.memcpy(hp, d, nbytes, 2); eng.device_free(d)
    cap = nbytes + nbytes // 8 + (1 << 20)
    ho = L.flacb200_host_alloc(cap)
    ptrs = (C.c_void_p * ntr)(*[hp + t * tb for t in range(ntr)])
    lens = (C.c_size_t * ntr)(*[tb] * ntr)
    out = np.zeros(16 * ntr, dtype=np.uint8)
    res = {"cores": os.cpu_count()}
    for th in (1, 2, 4, 8, 12, 15, 16):
        t0 = time.perf_counter()
        L.flacb200_md5_many(ptrs, lens, ntr, C.c_void_p(out.ctypes.data), th)
        res[f"md5_many_{th}_threads_gbs"] = round(nbytes / (time.perf_counter() - t0) / 1e9, 2)
    tracks = [((hp + t * tb, tb), n, rate, bps, ch, _abi.PCM_BYTES_LE) for t in range(ntr)]
    per = cap // ntr
    bufs = [(ho + t * per, per) for t in range(ntr)]
    encode_files(tracks, Options.best(), devices=[0], out_buffers=bufs)
    t0 = time.perf_counter()
    for _ in range(3):
        encode_files(tracks, Options.best(), devices=[0], out_buffers=bufs)
    res["encode_batch_ms"] = round((time.perf_counter() - t0) / 3 * 1e3, 1)
    print(json.dumps(res))

if __name__ == "__main__":
    main()
