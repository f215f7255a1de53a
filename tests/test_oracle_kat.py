"""Pins the CPU oracle against the reference's own known-answer tests and fixtures.

Every vector below is transcribed from a unit test / doc-test of the reference crate
(file:line cited per test) or is a fixture copied from its tests/data (tests/golden/ref_data).
"""
import ctypes as C
import hashlib

import numpy as np
import pytest

from oracle import oracle as fo
from flacb200_testutil import ref_file

f64 = np.float64


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


# ---- src/encode.rs:3503-3527 test_autocorrelation (exact equality) ----
@pytest.mark.parametrize(
    "windowed,order,expected",
    [
        ([1.0], 1, [1.0]),
        ([1.0, 2.0, 3.0, 4.0, 5.0], 4, [55.0, 40.0, 26.0, 14.0, 5.0]),
        (
            [0.0, 16.0, 31.0, 44.0, 54.0, 61.0, 64.0, 63.0, 58.0, 49.0, 38.0, 24.0, 8.0, -8.0, -24.0, -38.0, -49.0,
             -58.0, -63.0, -64.0, -61.0, -54.0, -44.0, -31.0, -16.0],
            4,
            [51408.0, 49792.0, 45304.0, 38466.0, 29914.0],
        ),
    ],
)
def test_autocorrelation(windowed, order, expected):
    w = np.array(windowed, dtype=f64)
    out = np.zeros(order + 1, dtype=f64)
    n = fo.lib().fo_autocorrelate(_dp(w), w.size, order, _dp(out))
    assert n == len(expected)
    assert out[:n].tolist() == expected


# ---- src/encode.rs:3591-3653 test_lp_coefficients_{1,2} (1e-6) ----
@pytest.mark.parametrize(
    "autoc,errors,coeffs",
    [
        (
            [55.0, 40.0, 26.0, 14.0, 5.0],
            [25.909091, 25.540351, 25.316142, 25.241623],
            [[0.727273], [0.814035, -0.119298], [0.802858, -0.043028, -0.093694],
             [0.797774, -0.045362, -0.050136, -0.054254]],
        ),
        (
            [51408.0, 49792.0, 45304.0, 38466.0, 29914.0],
            [3181.201369, 495.815931, 495.161449, 494.604514],
            [[0.968565], [1.858456, -0.918772], [1.891837, -0.986293, 0.036332],
             [1.890618, -0.953216, -0.027115, 0.033537]],
        ),
    ],
)
def test_lp_coefficients(autoc, errors, coeffs):
    r = np.array(autoc, dtype=f64)
    c = np.zeros(32 * 32, dtype=f64)
    e = np.zeros(32, dtype=f64)
    n = fo.lib().fo_lp_coefficients(_dp(r), r.size, _dp(c), _dp(e))
    assert n == 4
    for o in range(4):
        assert abs(e[o] - errors[o]) < 1e-6
        for j, v in enumerate(coeffs[o]):
            assert abs(c[o * 32 + j] - v) < 1e-6


# ---- src/encode.rs:3704-3745 test_compute_best_order (1e-6) ----
@pytest.mark.parametrize(
    "bps,precision,n,errors,expected",
    [
        (16, 5, 20, [3181.201369, 495.815931, 495.161449, 494.604514], [80.977565, 74.685594, 93.853530, 113.025628]),
        (16, 10, 4096, [15000.0, 25000.0, 20000.0, 30000.0], [1812.801817, 3346.934051, 2713.303385, 3935.492805]),
    ],
)
def test_subframe_bits_by_order(bps, precision, n, errors, expected):
    e = np.array(errors, dtype=f64)
    out = np.zeros(32, dtype=f64)
    cnt = fo.lib().fo_subframe_bits_by_order(bps, precision, n, _dp(e), e.size, _dp(out))
    assert cnt == 4
    for a, b in zip(out[:4], expected):
        assert abs(a - b) < 1e-6


# ---- src/encode.rs:3404-3476 test_quantization ----
def _quantize(coeffs, precision):
    c = np.array(coeffs, dtype=f64)
    q = np.zeros(32, dtype=np.int32)
    shift = C.c_uint32(0)
    rc = fo.lib().fo_quantize(c.size, _dp(c), precision, _ip(q), C.byref(shift))
    return rc, shift.value, q[: c.size].tolist()


def test_quantization():
    assert _quantize([0.797774, -0.045362, -0.050136, -0.054254], 10) == (0, 9, [408, -23, -25, -28])
    assert _quantize([-0.054687, -0.953216, -0.027115, 0.033537], 10) == (0, 9, [-28, -488, -14, 17])
    assert _quantize([0.0, 0.0, 0.0, 0.0], 10)[0] == 53  # ZeroLpCoefficients
    assert _quantize([-0.1, 0.1, 10000000.0, -0.2], 10) == (0, 0, [0, 0, 305, 0])
    assert _quantize([-0.1, 0.1, 100000000.0, -0.2], 10)[0] == 54  # LpNegativeShiftError


# ---- src/encode.rs:3216-3272 test_residual_encoding_{1,2} ----
@pytest.mark.parametrize(
    "samples,coefs,shift,expected",
    [
        (
            [0, 16, 31, 44, 54, 61, 64, 63, 58, 49, 38, 24, 8, -8, -24, -38, -49, -58, -63, -64, -61, -54, -44, -31, -16],
            [59, -30], 5,
            [2, 2, 2, 3, 3, 3, 2, 2, 3, 0, 0, 0, -1, -1, -1, -3, -2, -2, -2, -1, -1, 0, 0],
        ),
        (
            [64, 62, 56, 47, 34, 20, 4, -12, -27, -41, -52, -60, -63, -63, -60, -52, -41, -27, -12, 4, 20, 34, 47, 56, 62],
            [58, -29], 5,
            [2, 2, 0, 1, -1, -1, -1, -2, -2, -2, -1, -3, -2, 0, -1, 1, 0, 2, 2, 2, 4, 2, 4],
        ),
    ],
)
def test_residual_encoding(samples, coefs, shift, expected):
    x = np.array(samples, dtype=np.int32)
    q = np.array(coefs, dtype=np.int32)
    r = np.zeros(x.size, dtype=np.int32)
    rc = fo.lib().fo_lpc_residuals(q.size, shift, _ip(q), _ip(x), x.size, _ip(r))
    assert rc == 0
    assert r[: x.size - q.size].tolist() == expected


# ---- src/decode.rs:1754-1798 verify_prediction ----
@pytest.mark.parametrize(
    "coefs,shift,buf,expected",
    [
        ([-75, 166, 121, -269, -75, -399, 1042], 9,
         [-796, -547, -285, -32, 199, 443, 670, -2, -23, 14, 6, 3, -4, 12, -2, 10],
         [-796, -547, -285, -32, 199, 443, 670, 875, 1046, 1208, 1343, 1454, 1541, 1616, 1663, 1701]),
        ([119, -255, 555, -836, 879, -1199, 1757], 10,
         [-21363, -21951, -22649, -24364, -27297, -26870, -30017, 3157],
         [-21363, -21951, -22649, -24364, -27297, -26870, -30017, -29718]),
        ([709, -2589, 4600, -4612, 1350, 4220, -9743, 12671, -12129, 8586, -3775, -645, 3904, -5543, 4373, 182, -6873,
          13265, -15417, 11550], 12,
         [213238, 210830, 234493, 209515, 235139, 201836, 208151, 186277, 157720, 148176, 115037, 104836, 60794, 54523,
          412, 17943, -6025, -3713, 8373, 11764, 30094],
         [213238, 210830, 234493, 209515, 235139, 201836, 208151, 186277, 157720, 148176, 115037, 104836, 60794, 54523,
          412, 17943, -6025, -3713, 8373, 11764, 33931]),
    ],
)
def test_verify_prediction(coefs, shift, buf, expected):
    c = np.array(coefs[::-1], dtype=np.int64)  # coefficients.reverse()
    b = np.array(buf, dtype=np.int32)
    fo.lib().fo_predict(c.ctypes.data_as(C.POINTER(C.c_int64)), c.size, shift, _ip(b), b.size)
    assert b.tolist() == expected


# ---- src/stream.rs:1328-1356 frame number round trip (sampled) + UTF-8-like layout ----
def test_frame_number_roundtrip():
    L = fo.lib()
    buf = np.zeros(7, dtype=np.uint8)
    vals = list(range(0, 70000)) + [0x1FFFFF, 0x200000, 0x3FFFFFF, 0x4000000, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFFF]
    for v in vals:
        n = L.fo_write_frame_number(v, buf.ctypes.data_as(C.POINTER(C.c_uint8)))
        assert n > 0
        if v < (1 << 31) and v < 0x110000 and not (0xD800 <= v < 0xE000):
            assert bytes(buf[:n]) == chr(v).encode("utf-8")  # same layout as UTF-8 where UTF-8 is defined
        out = C.c_uint64(0)
        m = L.fo_read_frame_number(buf.ctypes.data_as(C.POINTER(C.c_uint8)), n, C.byref(out))
        assert (m, out.value) == (n, v)
    assert L.fo_write_frame_number(1 << 36, buf.ctypes.data_as(C.POINTER(C.c_uint8))) < 0


# ---- src/stream.rs:107-128 and :1645-1677: header bytes and the 12-byte CONSTANT frame incl. CRC-16 ----
HDR20 = bytes([0xFF, 0xF8, 0x69, 0x08, 0x00, 0x13, 0x64])


def test_frame_header_and_constant_frame_bytes():
    frame = HDR20 + bytes([0, 0, 0]) + bytes([0xD3, 0x3B])
    planar, h, used = fo.decode_frame(frame, None)
    assert used == 12
    assert (h.block_size, h.sample_rate, h.channels, h.bps, h.frame_number) == (20, 44100, 1, 16, 0)
    assert planar.tolist() == [[0] * 20]
    assert fo.crc8(HDR20[:-1]) == 0x64
    assert fo.crc16(frame[:-2]) == 0xD33B
    # and the encoder must emit exactly these bytes for 20 zero samples in a subset stream
    opt = fo.options("default")
    assert fo.encode_frame(opt, 44100, 16, np.zeros((1, 20), np.int32), 0, subset=True) == frame


def _frame_from_subframe(sub: bytes) -> bytes:
    body = HDR20 + sub
    return body + fo.crc16(body).to_bytes(2, "big")


# ---- src/stream.rs:2190-2223 FIXED-4 ----
def test_fixed_subframe_doc_kat():
    sub = bytes([0b0_001100_0, 0, 0, 0, 1, 0, 2, 0, 3, 0x00, 0x3F, 0xFF, 0xC0])
    planar, h, used = fo.decode_frame(_frame_from_subframe(sub), None)
    assert planar.tolist() == [list(range(20))]


# ---- src/stream.rs:2266-2311 LPC-1, precision 12, shift 11, coeff 1989, rice 1 ----
def test_lpc_subframe_doc_kat():
    sub = bytes([0b0_100000_0, 0x00, 0x00, 0b1011_0101, 0b1_0111110, 0b00101_000,
                 0x02, 0x88, 0x88, 0x88, 0x88, 0x88, 0x88, 0x88, 0x88, 0x88, 0x80])
    planar, h, used = fo.decode_frame(_frame_from_subframe(sub), None)
    x = [0]
    res = [1] + [2] * 18
    for r in res:
        x.append(r + ((1989 * x[-1]) >> 11))
    assert planar.tolist() == [x]


# ---- tests/data/all-frames.flac: CONSTANT, FIXED-4, LPC-1, VERBATIM; MD5 in STREAMINFO ----
def test_all_frames_fixture():
    flac = ref_file("all-frames.flac")
    pcm, si, md5 = fo.decode_stream(flac, want_md5=True)
    assert (si.sample_rate, si.bps, si.channels, si.total_samples) == (44100, 16, 1, 80)
    assert bytes(si.md5).hex() == "f53f86876dcd7783225c93ba8a938c7d"
    assert md5 == bytes(si.md5)
    assert pcm[:20].tolist() == [0] * 20
    assert pcm[20:40].tolist() == list(range(20))
    assert pcm[60:80].tolist() == list(range(20))


# ---- tests/seek.rs:10-31: sine.flac decodes to the STREAMINFO MD5 831671b8... ----
def test_sine_fixture_md5():
    flac = ref_file("sine.flac")
    pcm, si, md5 = fo.decode_stream(flac, want_md5=True)
    assert (si.sample_rate, si.bps, si.channels, si.total_samples, si.max_block_size) == (44100, 16, 2, 200000, 4096)
    assert md5.hex() == "831671b807f97051301e01d68b5c54b3"
    assert bytes(si.md5) == md5
    # cross-check our MD5 with hashlib
    assert hashlib.md5(fo.samples_to_bytes(pcm, 2)).digest() == md5


# ---- tests/data/cuesheet.flac: block 65535 (16-bit uncommon size), CONSTANT frames ----
def test_cuesheet_fixture_md5():
    flac = ref_file("cuesheet.flac")
    pcm, si, md5 = fo.decode_stream(flac, want_md5=True)
    assert (si.channels, si.bps, si.total_samples, si.max_block_size) == (2, 16, 48720504, 65535)
    assert md5.hex() == "2ae74d9f65a6acb8a4e9079125d68952"
    assert not pcm.any()


# ---- tests/corruption.rs:9-43: any single bit flip after the metadata must be detected ----
def test_corruption_detected():
    flac = bytearray(ref_file("sine.flac"))
    rng = np.random.default_rng(1234)
    for _ in range(100):
        pos = int(rng.integers(136, len(flac)))
        bit = 1 << int(rng.integers(0, 8))
        flac[pos] ^= bit
        with pytest.raises(fo.OracleError):
            fo.decode_stream(bytes(flac))
        flac[pos] ^= bit


# ---- src/byteorder.rs:188-243 sample (de)serialisation ----
@pytest.mark.parametrize("nbytes", [1, 2, 3, 4])
@pytest.mark.parametrize("be", [False, True])
def test_byteorder(nbytes, be):
    rng = np.random.default_rng(nbytes)
    lo, hi = -(1 << (8 * nbytes - 1)), (1 << (8 * nbytes - 1)) - 1
    s = rng.integers(lo, hi, size=4096, endpoint=True).astype(np.int32)
    s[:4] = [lo, hi, 0, -1]
    b = fo.samples_to_bytes(s, nbytes, be)
    expect = b"".join(int(v).to_bytes(nbytes, "big" if be else "little", signed=True) for v in s)
    assert b == expect
    assert fo.bytes_to_samples(b, nbytes, be).tolist() == s.tolist()


def test_md5_and_crc_selfcheck():
    data = bytes(range(256)) * 37
    assert fo.md5(data) == hashlib.md5(data).digest()
    assert fo.md5(b"") == hashlib.md5(b"").digest()
    # CRC-16/UMTS ("123456789") = 0xFEE8, CRC-8 (poly 7) = 0xF4: catalogue check values
    assert fo.crc16(b"123456789") == 0xFEE8
    assert fo.crc8(b"123456789") == 0xF4


# ---- Rice parameter: the integer form used on the GPU equals the reference's f64 form ----
def test_rice_parameter_integer_equivalence():
    L = fo.lib()
    rng = np.random.default_rng(7)

    def int_form(s, n):
        k = 0
        while (n << k) < s:
            k += 1
        return k

    for n in [1, 2, 3, 15, 16, 63, 64, 255, 256, 1000, 4080, 4096, 65535]:
        sums = set()
        for k in range(0, 34):
            for d in (-2, -1, 0, 1, 2):
                sums.add((n << k) + d)
        sums |= {int(v) for v in rng.integers(n + 1, n << 20, size=200)}
        for s in sums:
            if s > n and s < (1 << 48):
                assert L.fo_rice_parameter_f64(s, n) == int_form(s, n), (s, n)
