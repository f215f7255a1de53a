import os, sys, time, numpy as np
sys.path.insert(0, '.')
from flac_codec_b200 import Engine, Options, _abi
eng = Engine(0); eng.set_keep_info(False)
rate, bps, ch, ntr, n = 48000, 24, 2, int(os.environ.get("TRACKS","128")), 48000*180
nbytes = ntr*n*ch*3
d_pcm = eng.device_alloc(nbytes); eng.synth_pcm(d_pcm, 0, ntr, n, ch, rate, bps)
cap = nbytes + nbytes//8 + (1<<20)
d_out = eng.device_alloc(cap)
segs = [(t*n, n, 0) for t in range(ntr)]
_, sizes, total = eng.encode(Options.best(), rate, bps, ch, d_pcm, nbytes, _abi.PCM_BYTES_LE, segs, pcm_location=_abi.DEVICE, out=d_out, out_capacity=cap, out_location=_abi.DEVICE, want_sizes=True)
per=(n+4095)//4096
offs=np.concatenate([[0],np.cumsum(sizes.astype(np.int64))])
dsegs=[(int(offs[t*per]), int(offs[(t+1)*per]-offs[t*per]), t*n, n) for t in range(ntr)]
d_back=eng.device_alloc(nbytes)
eng.set_profiling(True)
for i in range(3):
    t0=time.perf_counter()
    nf,ns=eng.decode(rate,bps,ch,4096,d_out,total,dsegs,d_back,nbytes,_abi.PCM_BYTES_LE,frames_location=_abi.DEVICE,pcm_location=_abi.DEVICE)
    dt=time.perf_counter()-t0
    tm=eng.timings()
    print(nf, ns, round(dt*1e3,2), tm.launches, [round(x,2) for x in tm.kernel_ms[:5]], round(tm.total_ms,2))
