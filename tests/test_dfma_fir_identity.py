"""The two bit-level identities k_analyze3's FIR rests on (flac_codec_b200/csrc/encode_analyze.cu, FLACB200_A3_FIR = 1), restated with
numpy float64 (every value involved is an integer below 2^53, so IEEE products and sums are exact with or without fusing):

  * int32 -> double without the conversion unit: the double whose words are (0x43300000, x ^ 0x80000000) is 2^52 + 2^31 + x, and
    subtracting 2^52 + 2^31 gives x exactly;
  * the i64 dot product of the reference (src/encode.rs:3187) accumulated on top of 1.5 * 2^52 leaves its two's complement bits in
    the mantissa: `(sum >> shift) as i32` is a funnel shift over the two words of the double for shift <= 15."""
import numpy as np


def test_int_to_double_by_bit_pasting():
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.integers(-(1 << 31), 1 << 31, 100000), [0, 1, -1, (1 << 31) - 1, -(1 << 31)]]).astype(np.int64)
    words = (np.uint64(0x43300000) << np.uint64(32)) | ((x.astype(np.uint64) ^ np.uint64(0x80000000)) & np.uint64(0xFFFFFFFF))
    d = words.view(np.float64) - 4503601774854144.0
    assert np.array_equal(d, x.astype(np.float64))


def test_biased_accumulator_holds_the_integer_sum():
    rng = np.random.default_rng(2)
    for trial in range(3000):
        taps = int(rng.integers(1, 17))
        qbits, xbits = int(rng.integers(2, 16)), int(rng.integers(2, 26))
        q = rng.integers(-(1 << (qbits - 1)), 1 << (qbits - 1), taps).astype(np.int64)
        x = rng.integers(-(1 << (xbits - 1)), 1 << (xbits - 1), taps).astype(np.int64)
        if trial % 7 == 0:   # extremes
            q[:] = -(1 << 14)
            x[:] = -(1 << 24) if trial % 2 else (1 << 24) - 1
        acc = np.float64(6755399441055744.0)   # 1.5 * 2^52
        for j in range(taps):
            acc = np.float64(q[j]) * np.float64(x[j]) + acc
        total = int((q * x).sum())
        bits = int(np.array([acc]).view(np.uint64)[0])
        lo, hi = bits & 0xFFFFFFFF, bits >> 32
        for shift in (0, 1, 7, 15):
            got = ((hi << 32 | lo) >> shift) & 0xFFFFFFFF   # __funnelshift_r(lo, hi, shift)
            want = (total >> shift) & 0xFFFFFFFF            # (sum >> shift) as i32, as unsigned bits
            assert got == want, (trial, shift, total)
