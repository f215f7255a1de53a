"""Shared helpers for the test-suite (fixtures, deterministic signal generators)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_DATA = os.path.join(GOLDEN, "ref_data")


def ref_file(name: str) -> bytes:
    with open(os.path.join(REF_DATA, name), "rb") as f:
        return f.read()


# ---- sine generators of the reference's tests/format.rs:687-774 (including their precedence quirk:
# only the second term is scaled by full_scale) ----
def generate_sine_1(full_scale, sample_rate, samples, f1, a1, f2, a2):
    delta1 = 2.0 * np.pi / (sample_rate / f1)
    delta2 = 2.0 * np.pi / (sample_rate / f2)
    k = np.arange(samples, dtype=np.float64)
    t1, t2 = k * delta1, k * delta2
    val = a1 * np.sin(t1) + a2 * np.sin(t2) * full_scale
    return np.trunc(val).astype(np.int64).clip(-(1 << 31), (1 << 31) - 1).astype(np.int32)


def generate_sine_2(full_scale, sample_rate, samples, f1, a1, f2, a2, fmult):
    delta1 = 2.0 * np.pi / (sample_rate / f1)
    delta2 = 2.0 * np.pi / (sample_rate / f2)
    k = np.arange(samples, dtype=np.float64)
    t1, t2 = k * delta1, k * delta2
    c0 = a1 * np.sin(t1) + a2 * np.sin(t2) * full_scale
    c1 = -(a1 * np.sin(t1 * fmult)) + a2 * np.sin(t2 * fmult) * full_scale
    out = np.empty(samples * 2, dtype=np.float64)
    out[0::2], out[1::2] = c0, c1
    return np.trunc(out).astype(np.int64).clip(-(1 << 31), (1 << 31) - 1).astype(np.int32)


# ---- deterministic integer-only synthetic PCM: mixed sinusoids, a chirp and
# noise, a silence gap and a full-scale square burst.  The product ships the same generator as a CUDA
# kernel (flacb200_synth_pcm); this numpy statement is what the tests compare it with. ----
SINE_LUT_BITS = 12


def sine_lut():
    k = np.arange(1 << SINE_LUT_BITS, dtype=np.float64)
    return np.round(np.sin(2.0 * np.pi * k / (1 << SINE_LUT_BITS)) * ((1 << 30) - 1)).astype(np.int64)


def splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def synth_pcm(track, channels, n, sample_rate, bps, seed=20261017):
    """Returns int32 [n, channels] interleaved-ready PCM, bit-identical to the CUDA generator."""
    lut = sine_lut()
    out = np.zeros((n, channels), dtype=np.int32)
    idx = np.arange(n, dtype=np.uint64)
    old = np.seterr(over="ignore")
    try:
        full = np.int64(1) << np.int64(bps - 1)
        base = None
        for c in range(channels):
            cid = np.uint64(track * 8 + c + 1)
            freqs = [220.0, 440.0 * (1.0 + c / 16.0), 3520.0]
            amps = [0.30, 0.20, 0.05]
            acc = np.zeros(n, dtype=np.int64)
            for f, a in zip(freqs, amps):
                delta = np.uint64(int(f / sample_rate * (1 << 32)))
                phase0 = np.uint64((track * 977 + c * 131) << 20) & np.uint64(0xFFFFFFFF)
                ph = (phase0 + idx * delta) & np.uint64(0xFFFFFFFF)
                s = lut[(ph >> np.uint64(32 - SINE_LUT_BITS)).astype(np.int64)]
                acc += (s * np.int64(int(a * 1024))) >> np.int64(10)
            # chirp 100 Hz -> 8 kHz over 2^22 samples (phase quadratic in n, u64 wraparound)
            d0 = np.uint64(int(100.0 / sample_rate * (1 << 32)))
            dd = np.uint64(int((8000.0 - 100.0) / sample_rate * (1 << 32) / (1 << 22)))
            ph = (idx * d0 + ((idx * idx) >> np.uint64(1)) * dd) & np.uint64(0xFFFFFFFF)
            acc += (lut[(ph >> np.uint64(32 - SINE_LUT_BITS)).astype(np.int64)] * np.int64(154)) >> np.int64(10)  # 0.15
            # scale from 31-bit to bps
            sig = (acc >> np.int64(31 - bps)) if bps <= 31 else (acc << np.int64(bps - 31))
            if c == 1 and base is not None:
                sig = (base * np.int64(819)) >> np.int64(10)  # 0.8 * channel 0
            if c == 0:
                base = sig.copy()
            # noise: ~ -48 dBFS at 16 bit, -72 at 24, -96 at 32 -> 8 noise bits for all widths
            r = splitmix64(np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15) * cid + idx)
            noise = (r >> np.uint64(56)).astype(np.int64) - np.int64(128)
            sig = sig + noise
            # silence gap [0.5s, 1.0s) and full-scale square burst [1.5s, 1.5s + 4096)
            gap = (idx >= np.uint64(sample_rate // 2)) & (idx < np.uint64(sample_rate))
            sig = np.where(gap, 0, sig)
            b0 = np.uint64(sample_rate * 3 // 2)
            burst = (idx >= b0) & (idx < b0 + np.uint64(4096))
            sq = np.where(((idx >> np.uint64(5)) & np.uint64(1)) == 0, full - 1, -full)
            sig = np.where(burst, sq, sig)
            out[:, c] = np.clip(sig, -full, full - 1).astype(np.int32)
    finally:
        np.seterr(**old)
    return out
