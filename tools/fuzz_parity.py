#!/usr/bin/env python
"""Randomised differential test: GPU encode must be byte-identical to the oracle's frames, GPU decode (both decoders) of
those frames bit-exact, over random stream shapes (channels, bits, block size, preset, LPC order, partition order, length,
signal kind, wasted bits).  Diagnostic tool (executes oracle/): `python tools/fuzz_parity.py [iterations] [seed]`."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from flac_codec_b200 import Engine, Options, _abi  # noqa: E402
from flacb200_testutil import synth_pcm  # noqa: E402
from oracle import oracle as fo  # noqa: E402


def signal(rng, kind, ch, n, rate, bps):
    if kind == 0:
        x = synth_pcm(int(rng.integers(0, 1000)), ch, n, rate, bps)
    elif kind == 1:   # noise at a random level
        lvl = int(rng.integers(1, bps))
        x = rng.integers(-(1 << (lvl - 1)), 1 << (lvl - 1), size=(n, ch), dtype=np.int64).astype(np.int32)
    elif kind == 2:   # full-scale square / constant stretches
        x = np.where((np.arange(n)[:, None] // int(rng.integers(1, 300))) % 2 == 0, (1 << (bps - 1)) - 1, -(1 << (bps - 1))).astype(np.int32)
        x = np.repeat(x, ch, axis=1)[:, :ch]
        x[: n // 3] = int(rng.integers(-5, 5))
    else:             # slow ramp + tiny noise (low orders win)
        x = (np.arange(n)[:, None] * int(rng.integers(1, 50)) % (1 << (bps - 2))).astype(np.int32) + rng.integers(-2, 3, size=(n, ch)).astype(np.int32)
    w = int(rng.integers(0, 4)) if rng.random() < 0.2 else 0   # wasted bits
    lo, hi = -(1 << (bps - 1)), (1 << (bps - 1)) - 1
    return np.clip((x >> w) << w, lo, hi).astype(np.int32)


def run(iters, seed, eng=None):
    """Returns None, or a description of the first mismatch."""
    rng = np.random.default_rng(seed)
    eng = eng if eng is not None else Engine(0)
    for it in range(iters):
        ch = int(rng.choice([1, 2, 2, 2, 3, 6, 8]))
        bps = int(rng.choice([8, 12, 16, 16, 20, 24, 24, 32]))
        rate = int(rng.choice([8000, 44100, 48000, 96000, 192000]))
        preset = str(rng.choice(["fast", "default", "best"]))
        block = int(rng.choice([16, 192, 576, 1152, 4096, 4096, 4608, int(rng.integers(16, 8192))]))
        n = int(rng.integers(1, 6 * block + 50))
        kw = {"block_size": block}
        o = getattr(Options, preset)().block_size(block)
        if rng.random() < 0.4:
            order = int(rng.integers(1, 33))
            kw["max_lpc_order"] = order
            o = o.max_lpc_order(order)
        if rng.random() < 0.3:
            po = int(rng.integers(0, 7))
            kw["max_partition_order"] = po
            o = o.max_partition_order(po)
        x = signal(rng, int(rng.integers(0, 4)), ch, n, rate, bps)
        desc = f"it={it} ch={ch} bps={bps} rate={rate} preset={preset} {kw} n={n}"
        ref, ref_sizes = fo.encode_frames_only(fo.options(preset, **kw), rate, bps, ch, x.reshape(-1))
        nb = (bps + 7) // 8
        raw = np.frombuffer(fo.samples_to_bytes(x.reshape(-1), nb), dtype=np.uint8).copy()
        data, sizes, total = eng.encode(o, rate, bps, ch, raw, raw.nbytes, _abi.PCM_BYTES_LE, [(0, n, 0)])
        if data.tobytes() != ref:
            return f"ENCODE MISMATCH {desc} ({total} vs {len(ref)} bytes)"
        # what the reference's own decoder makes of these frames: normally the input, but its encoder can emit frames its
        # decoder rejects (e.g. a 4-sample block with a fixed order 2 subframe at partition order 1: the first partition is
        # empty, `rchunks` then yields one partition instead of two -> InvalidPartitionOrder, src/decode.rs:1816-1820)
        want_err = None
        off = 0
        for fi, sz in enumerate(ref_sizes):
            try:
                fo.decode_frame(ref[off:off + int(sz)], None, 0)
            except fo.OracleError as oe:
                want_err = (oe.code, fi)
                break
            off += int(sz)
        packed = bps <= 24 and ch <= 2   # packed little-endian bytes: k_restore_emit (default) and k_restore + k_emit4 (bit 512)
        for legacy, kind in ((0, _abi.PCM_I32_INTERLEAVED), (64, _abi.PCM_I32_INTERLEAVED)) + (((0, _abi.PCM_BYTES_LE), (512, _abi.PCM_BYTES_LE)) if packed else ()):
            label = {0: "k_parse+k_restore", 64: "k_decode", 512: "unfused"}[legacy] + (" bytes" if kind == _abi.PCM_BYTES_LE else "")
            eng.set_option("legacy", legacy)
            out = np.zeros(x.size, dtype=np.int32) if kind == _abi.PCM_I32_INTERLEAVED else np.zeros(x.size * nb, dtype=np.uint8)
            buf = np.frombuffer(ref, dtype=np.uint8).copy()
            try:
                nf, ns = eng.decode(rate, bps, ch, block, buf, buf.size, [(0, buf.size, 0, n)], out, out.nbytes, kind)
            except _abi.FlacB200Error as e:
                if want_err == (e.code, e.bad_frame):
                    continue   # same error at the same frame as the reference decoder
                eng.set_option("legacy", 0)
                return f"DECODE ERROR {desc} ({label}): {e}; the oracle: {want_err}"
            if want_err is not None:
                eng.set_option("legacy", 0)
                return f"DECODE ACCEPTED what the oracle rejects {desc} ({label}): {want_err}"
            got = out if kind == _abi.PCM_I32_INTERLEAVED else fo.bytes_to_samples(out.tobytes(), nb)
            if ns != n or not np.array_equal(np.asarray(got).reshape(-1, ch), x):
                eng.set_option("legacy", 0)
                return f"DECODE MISMATCH {desc} ({label}, {ns} samples)"
        eng.set_option("legacy", 0)
    return None


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    t0 = time.time()
    bad = run(iters, seed)
    print(bad if bad else f"fuzz ok: {iters} streams, seed {seed}, {time.time() - t0:.1f} s")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
