#!/usr/bin/env python
"""Single-bit corruptions of a .flac file: the GPU decoders' verdict (error ordinal, failing frame index, samples before it)
against the oracle's serial reader.  Prints one JSON line per disagreement and a summary line.
usage: corrupt_probe.py [file] [trials] [seed]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from flac_codec_b200 import Engine, _abi  # noqa: E402
from oracle import oracle as fo  # noqa: E402


def frame_offsets(flac, si):
    offs, p, cur = [], si.frames_start, 0
    while cur < si.total_samples:
        _, h, used = fo.decode_frame(flac[p:], si, si.total_samples - cur)
        offs.append(p)
        p += used
        cur += h.block_size
    offs.append(p)
    return offs


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "ref_data", "sine.flac")
    trials = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 1234
    if path.startswith("synth:"):   # synth:<bps>:<channels>:<preset>:<pcm frames>[:<block size>]
        from flacb200_testutil import synth_pcm
        f = path.split(":")
        bps, nch, preset, n = int(f[1]), int(f[2]), f[3], int(f[4])
        kw = {"block_size": int(f[5])} if len(f) > 5 else {}
        rate = 48000
        x = synth_pcm(11, nch, n, rate, bps)
        flac = bytearray(fo.encode_stream(fo.options(preset, **kw), rate, bps, nch, x.reshape(-1))[0])
    else:
        flac = bytearray(open(path, "rb").read())
    si = fo.read_streaminfo(bytes(flac))
    offs = frame_offsets(bytes(flac), si)
    eng = Engine(0)
    rng = np.random.default_rng(seed)
    ch = si.channels
    bad = 0
    for t in range(trials):
        pos = int(rng.integers(si.frames_start * 8, len(flac) * 8))
        flac[pos >> 3] ^= 0x80 >> (pos & 7)
        code, nf, ns, ref = fo.decode_stream_ex(bytes(flac))
        fr = int(np.searchsorted(offs, pos >> 3, side="right")) - 1
        for legacy in (0, 64):
            eng.set_option("legacy", legacy)
            frames = np.frombuffer(bytes(flac), dtype=np.uint8)[si.frames_start:].copy()
            out = np.zeros(si.total_samples * ch, dtype=np.int32)
            g_code, g_bad = 0, -1
            try:
                eng.decode(si.sample_rate, si.bps, ch, si.max_block_size, frames, frames.size, [(0, frames.size, 0, si.total_samples)],
                           out, out.nbytes, _abi.PCM_I32_INTERLEAVED)
            except _abi.FlacB200Error as e:
                g_code, g_bad = e.code, e.bad_frame
            prefix_ok = bool(np.array_equal(out[:ns], ref))
            if g_code != code or (code and g_bad != nf) or not prefix_ok:
                bad += 1
                print(json.dumps({"trial": t, "bit": pos, "byte_in_frame": (pos >> 3) - offs[fr], "frame": fr, "frame_len": offs[fr + 1] - offs[fr],
                                  "oracle": [code, nf, ns], "gpu": [g_code, g_bad], "prefix_ok": prefix_ok, "legacy": legacy}))
        eng.set_option("legacy", 0)
        flac[pos >> 3] ^= 0x80 >> (pos & 7)
    print(json.dumps({"file": os.path.basename(path), "trials": trials, "disagreements": bad}))


if __name__ == "__main__":
    main()
