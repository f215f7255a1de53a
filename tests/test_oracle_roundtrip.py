"""The reference's integration matrix (tests/format.rs) replayed against the CPU oracle:
encode -> decode must be lossless, and the container fields must be self-consistent.
The decoder half of the oracle is pinned independently by libFLAC-made fixtures
(test_oracle_kat.py), so a symmetric encoder/decoder bug cannot hide here."""
import hashlib
import os

import numpy as np
import pytest

from flacb200_testutil import REF_DATA, generate_sine_1, generate_sine_2, ref_file, synth_pcm
from oracle import oracle as fo


def roundtrip(opt, rate, bps, channels, samples, total_known=True):
    flac, sizes = fo.encode_stream(opt, rate, bps, channels, samples, total_known=total_known)
    pcm, si, md5 = fo.decode_stream(flac, want_md5=True)
    assert pcm.tolist() == np.asarray(samples).reshape(-1).tolist()
    assert md5 == bytes(si.md5)
    assert si.total_samples == len(pcm) // channels
    if len(sizes):
        assert si.min_frame_size == sizes.min() and si.max_frame_size == sizes.max()
    return flac, sizes


# tests/format.rs:17 test_small_files
@pytest.mark.parametrize(
    "channels,data",
    [
        (1, b"\x00\x80"),
        (2, b"\x00\x80\xff\x7f"),
        (1, b"\xe7\xff\x00\x00\x19\x00\x32\x00\x64\x00"),
        (2, b"\xe7\xff\xf4\x01\x00\x00\x90\x01\x19\x00\x2c\x01\x32\x00\xc8\x00\x64\x00\x64\x00"),
    ],
)
def test_small_files(channels, data):
    opt = fo.options("fast", max_lpc_order=16, mid_side=1, padding=None)
    roundtrip(opt, 44100, 16, channels, fo.bytes_to_samples(data, 2))


# tests/format.rs:85 test_blocksize_variations
def test_blocksize_variations():
    data = fo.bytes_to_samples(ref_file("noise32.raw"), 1)
    for blocksize in range(16, 34):
        for lpc_order in [0, 1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 31, 32]:
            opt = fo.options("best", max_lpc_order=lpc_order or None, block_size=blocksize, padding=None)
            roundtrip(opt, 44100, 8, 1, data)


# tests/format.rs:137 test_fractional (noise.raw is 1.5 MiB of random bytes; any random bytes do)
def test_fractional():
    rng = np.random.default_rng(42)
    noise = rng.integers(-32768, 32767, size=16390 * 2, endpoint=True).astype(np.int32)
    cases = [(33, [31, 32, 33, 34, 35, 2046, 2047, 2048, 2049, 2050]),
             (256, [254, 255, 256, 257, 258, 510, 511, 512, 513, 514, 1022, 1023, 1024, 1025, 1026, 2046, 2047, 2048,
                    2049, 2050, 4094, 4095, 4096, 4097, 4098]),
             (2048, [1022, 1023, 1024, 1025, 1026, 2046, 2047, 2048, 2049, 2050, 4094, 4095, 4096, 4097, 4098]),
             (4608, [1022, 1023, 1024, 1025, 1026, 2046, 2047, 2048, 2049, 2050, 4094, 4095, 4096, 4097, 4098, 4606,
                     4607, 4608, 4609, 4610, 8190, 8191, 8192, 8193, 8194, 16382, 16383, 16384, 16385, 16386])]
    for blocksize, counts in cases:
        opt = fo.options("default", block_size=blocksize, padding=None)
        for samples in counts:
            roundtrip(opt, 44100, 16, 2, noise[: samples * 2])


# tests/format.rs:208 test_roundtrip (36 fixtures)
@pytest.mark.parametrize("channels", [1, 2, 4, 8])
@pytest.mark.parametrize("bps", [8, 16, 24])
@pytest.mark.parametrize("frames", [1, 111, 4777])
def test_roundtrip_fixtures(channels, bps, frames):
    raw = ref_file(f"roundtrip-{channels}-{bps}-{frames}.raw")
    samples = fo.bytes_to_samples(raw, bps // 8)
    for preset in ("default", "fast", "best"):
        roundtrip(fo.options(preset), 44100, bps, channels, samples)
    roundtrip(fo.options("default"), 44100, bps, channels, samples, total_known=False)


# tests/format.rs:438 test_full_scale_deflection
@pytest.mark.parametrize("bps", [8, 16, 24, 32])
def test_full_scale_deflection(bps):
    hi, lo = (1 << (bps - 1)) - 1, -(1 << (bps - 1))
    patterns = [[hi] * 2, [lo] * 2, [hi, lo], [lo, hi], [hi, hi, lo], [lo, lo, hi], [hi, lo, lo], [lo, hi, hi],
                [hi, hi, lo, lo], [hi, lo, hi, hi, lo, lo, hi]]
    for pat in patterns:
        x = np.array((pat * 1200)[:4096 + 37], dtype=np.int32)
        for preset in ("default", "best"):
            roundtrip(fo.options(preset), 44100, bps, 1, x)
            roundtrip(fo.options(preset), 44100, bps, 2, np.concatenate([x, x])[: 2 * (len(x) // 2) * 2 // 2 * 1])


# tests/format.rs:624 test_wasted_bits (asserts wasted_bps > 0 in the subframe)
def test_wasted_bits():
    x = fo.bytes_to_samples(ref_file("wasted-bits.raw"), 2)
    assert np.bitwise_or.reduce(x) == 8188
    opt = fo.options("default")
    roundtrip(opt, 44100, 16, 1, x)
    _, info = fo.encode_frame(opt, 44100, 16, x[None, :2000], want_info=True)
    assert info.sub[0].wasted == 2


# tests/format.rs:777 test_sine_wave_streams (subset of the 20 recipes x 4 widths)
@pytest.mark.parametrize("bps", [8, 16, 24, 32])
def test_sine_streams(bps):
    fs = float((1 << (bps - 1)) - 1)
    mono = [(441.0, 0.50, 441.0, 0.49), (441.0, 0.61, 661.5, 0.37), (441.0, 0.50, 882.0, 0.49),
            (8820.0, 0.70, 4410.0, 0.29)]
    for f1, a1, f2, a2 in mono:
        x = generate_sine_1(fs, 48000.0, 20000, f1, a1, f2, a2)
        roundtrip(fo.options("default"), 48000, bps, 1, x)
    stereo = [(441.0, 0.50, 441.0, 0.49, 1.0), (441.0, 0.61, 661.5, 0.37, 2.0), (8820.0, 0.70, 4410.0, 0.29, 0.5)]
    for f1, a1, f2, a2, fm in stereo:
        x = generate_sine_2(fs, 44100.0, 20000, f1, a1, f2, a2, fm)
        for preset in ("default", "best", "fast"):
            roundtrip(fo.options(preset), 44100, bps, 2, x)


# tests/format.rs:1248-1384 test_noise_* (reduced sizes)
@pytest.mark.parametrize("bps", [8, 16, 24, 32])
@pytest.mark.parametrize("channels", [1, 2, 4, 8])
def test_noise(bps, channels):
    rng = np.random.default_rng(bps * 10 + channels)
    lo, hi = -(1 << (bps - 1)), (1 << (bps - 1)) - 1
    n = 70000
    x = rng.integers(lo, hi, size=n * channels, endpoint=True).astype(np.int64).astype(np.int32)
    for preset, bs in (("default", 4096), ("fast", 32), ("best", 32768), ("default", 65535)):
        m = n if bs >= 4096 else 1000
        roundtrip(fo.options(preset, block_size=bs), 44100, bps, channels, x[: m * channels])


def test_synthetic_bench_signal_compresses_and_roundtrips():
    x = synth_pcm(0, 2, 3 * 44100, 44100, 16)
    flac, sizes = roundtrip(fo.options("default"), 44100, 16, 2, x.reshape(-1))
    ratio = len(flac) / (x.size * 2)
    assert 0.2 < ratio < 0.9
    x = synth_pcm(3, 2, 2 * 48000, 48000, 24)
    flac, sizes = roundtrip(fo.options("best"), 48000, 24, 2, x.reshape(-1))
    assert 0.2 < len(flac) / (x.size * 3) < 0.9


def test_multithreaded_encode_is_identical():
    x = synth_pcm(1, 2, 100000, 44100, 16).reshape(-1)
    a, sa = fo.encode_stream(fo.options("default"), 44100, 16, 2, x, nthreads=1)
    b, sb = fo.encode_stream(fo.options("default"), 44100, 16, 2, x, nthreads=4)
    assert a == b and sa.tolist() == sb.tolist()


def test_seektable_and_streaminfo_layout():
    # 25 s of 8 kHz mono, block 4096, default seektable every 10 s -> 3 points
    x = (np.arange(25 * 8000) % 251).astype(np.int32)
    flac, sizes = fo.encode_stream(fo.options("default"), 8000, 16, 1, x)
    assert flac[:4] == b"fLaC"
    assert flac[4] == 0x00 and int.from_bytes(flac[5:8], "big") == 34
    p = 4 + 4 + 34
    assert flac[p] == 0x03 and int.from_bytes(flac[p + 1:p + 4], "big") == 3 * 18
    offs = np.concatenate([[0], np.cumsum(sizes)])
    for k in range(3):
        q = p + 4 + 18 * k
        so = int.from_bytes(flac[q:q + 8], "big")
        bo = int.from_bytes(flac[q + 8:q + 16], "big")
        fs = int.from_bytes(flac[q + 16:q + 18], "big")
        target = k * 10 * 8000
        assert so <= target < so + fs
        assert bo == offs[so // 4096]
    p += 4 + 3 * 18
    assert flac[p] == 0x81 and int.from_bytes(flac[p + 1:p + 4], "big") == 4096
    si = fo.read_streaminfo(flac)
    assert si.frames_start == p + 4 + 4096
    # unknown length: SEEKTABLE is carved out of PADDING and placed after it
    flac2, _ = fo.encode_stream(fo.options("default"), 8000, 16, 1, x, total_known=False)
    assert len(flac2) == len(flac) - (4 + 3 * 18)
    q = 4 + 4 + 34
    assert flac2[q] == 0x01 and int.from_bytes(flac2[q + 1:q + 4], "big") == 4096 - (4 + 3 * 18)
    pcm, si2 = fo.decode_stream(flac2)
    assert pcm.tolist() == x.tolist()
