"""GPU-vs-oracle parity probe: encodes several configurations through the C ABI and prints the first
differences (per-subframe decisions) -- a debugging aid, also used by tests/test_gpu_encode.py."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from flac_codec_b200 import Engine, Options  # noqa: E402
from flac_codec_b200 import _abi  # noqa: E402


def sub_tuple(s):
    t = (s.type, s.order, s.wasted, s.bps, s.bits)
    if s.type == 3:
        t += (s.precision, s.shift, tuple(s.coefs[: s.order]))
    if s.type >= 2:
        n = 1 << s.partition_order
        t += (s.coding_method, s.partition_order, tuple(s.rice[:n]), tuple(s.kind[:n]))
    return t


def oracle_options(fo, opt: Options):
    o = opt.c
    return fo.options("default", block_size=o.block_size, max_lpc_order=o.max_lpc_order,
                      max_partition_order=o.max_partition_order, mid_side=o.mid_side,
                      exhaustive_channel_correlation=o.exhaustive_channel_correlation, window_kind=o.window_kind,
                      tukey_p=o.tukey_p)


def compare(eng: Engine, opt: Options, rate: int, bps: int, channels: int, interleaved: np.ndarray, label: str = "",
            first_frame_number: int = 0, verbose: bool = True, pcm_kind=None):
    """Returns (identical_frames, n_frames, gpu_bytes, oracle_bytes)."""
    from oracle import oracle as fo

    x = np.ascontiguousarray(interleaved, dtype=np.int32).reshape(-1)
    n_pcm = x.size // channels
    bytes_per_sample = (bps + 7) // 8
    if pcm_kind is None:
        pcm_kind = _abi.PCM_BYTES_LE
    if pcm_kind == _abi.PCM_BYTES_LE:
        raw = np.frombuffer(fo.samples_to_bytes(x, bytes_per_sample), dtype=np.uint8).copy()
    elif pcm_kind == _abi.PCM_BYTES_BE:
        raw = np.frombuffer(fo.samples_to_bytes(x, bytes_per_sample, True), dtype=np.uint8).copy()
    else:
        raw = x.copy()
    data, sizes, total = eng.encode(opt, rate, bps, channels, raw, raw.nbytes, pcm_kind, [(0, n_pcm, first_frame_number)])
    gpu = data.tobytes()
    ref, ref_sizes, infos = fo.encode_frames_only(oracle_options(fo, opt), rate, bps, channels, x, first_frame_number,
                                                  nthreads=4, want_infos=True)
    ginfos, gn = eng.last_info()
    same = 0
    goff = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    roff = np.concatenate([[0], np.cumsum(ref_sizes)]).astype(np.int64)
    shown = 0
    for f in range(len(ref_sizes)):
        g = gpu[goff[f]:goff[f + 1]] if f < len(sizes) else b""
        r = ref[roff[f]:roff[f + 1]]
        if g == r:
            same += 1
            continue
        if verbose and shown < 3:
            shown += 1
            print(f"[{label}] frame {f}: gpu {len(g)} B vs oracle {len(r)} B; assignment gpu "
                  f"{ginfos[f].channel_assignment if f < gn else '?'} oracle {infos[f].channel_assignment}")
            if f < gn:
                for k in range(infos[f].channels):
                    a, b = sub_tuple(ginfos[f].sub[k]), sub_tuple(infos[f].sub[k])
                    if a != b:
                        print(f"   sub {k}: gpu    {a}\n          oracle {b}")
                    else:
                        print(f"   sub {k}: decisions identical {a[:5]}")
            first = next((i for i in range(min(len(g), len(r))) if g[i] != r[i]), min(len(g), len(r)))
            print(f"   first differing byte {first}: gpu {g[first:first+8].hex()} oracle {r[first:first+8].hex()}")
    if verbose:
        print(f"[{label}] identical frames {same}/{len(ref_sizes)}  gpu {len(gpu)} B oracle {len(ref)} B")
    return same, len(ref_sizes), gpu, ref


def main():
    from flacb200_testutil import synth_pcm

    eng = Engine(0)
    cases = [
        ("16b stereo default", Options.default(), 44100, 16, 2, synth_pcm(0, 2, 44100 * 2 + 100, 44100, 16)),
        ("24b stereo best", Options.best(), 48000, 24, 2, synth_pcm(1, 2, 48000 * 2 + 77, 48000, 24)),
        ("16b mono default", Options.default(), 44100, 16, 1, synth_pcm(2, 1, 50000, 44100, 16)),
        ("16b stereo fast", Options.fast(), 44100, 16, 2, synth_pcm(3, 2, 50000, 44100, 16)),
        ("24b 8ch best", Options.best(), 96000, 24, 8, synth_pcm(4, 8, 20000, 96000, 24)),
        ("32b stereo best lpc32", Options.best().max_lpc_order(32), 192000, 32, 2, synth_pcm(5, 2, 30000, 192000, 32)),
    ]
    ok = True
    for label, opt, rate, bps, ch, x in cases:
        same, n, _, _ = compare(eng, opt, rate, bps, ch, x.reshape(-1), label)
        ok &= same == n
    print("ALL IDENTICAL" if ok else "MISMATCHES")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
