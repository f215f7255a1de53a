"""GPU decode parity (through the C ABI): decoded PCM must be bit-exact against the reference
decoder's behaviour -- the reference's own golden streams (tests/data/*.flac with their STREAMINFO
MD5), the CPU oracle on oracle-encoded streams, hand-built frames for the paths no encoder here emits
(33-bit side channel, a sync code inside the payload), and the corruption test of tests/corruption.rs."""
import hashlib

import numpy as np
import pytest

from flacb200_testutil import ref_file, synth_pcm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from flac_codec_b200 import Engine

    e = Engine(0)
    yield e
    e.close()


@pytest.fixture(autouse=True, params=["parse+restore", "thread-per-frame", "unfused"])
def decode_path(request, eng):
    """Every test runs over both decoders: k_parse + k_restore (default; packed 16/24-bit mono/stereo output goes through the
    fused k_restore_emit) and k_decode ("legacy" bit 64); "unfused" (bit 512) keeps k_restore + k_emit4 for those layouts."""
    eng.set_option("legacy", {"thread-per-frame": 64, "unfused": 512}.get(request.param, 0))
    yield request.param
    eng.set_option("legacy", 0)


@pytest.fixture(scope="module")
def fo():
    from oracle import oracle

    return oracle


def gpu_decode_stream(eng, flac: bytes, fo, kind=None, chunk=0, total_override=None):
    """Decode a whole .flac file's frames on the GPU; returns interleaved int32 samples."""
    from flac_codec_b200 import _abi

    si = fo.read_streaminfo(flac)
    frames = np.frombuffer(flac, dtype=np.uint8)[si.frames_start:].copy()
    total = si.total_samples if total_override is None else total_override
    ch, bps = si.channels, si.bps
    cap = total if total else len(frames) * 16
    kind = _abi.PCM_I32_INTERLEAVED if kind is None else kind
    bytes_per = 4 if kind >= 2 else (bps + 7) // 8
    out = np.zeros(cap * ch * bytes_per, dtype=np.uint8)
    eng.set_chunk_frames(chunk)
    try:
        nf, ns = eng.decode(si.sample_rate, bps, ch, si.max_block_size, frames, frames.size, [(0, frames.size, 0, total)], out,
                            out.nbytes, kind, planar_stride=cap if kind == _abi.PCM_I32_PLANAR else 0)
    finally:
        eng.set_chunk_frames(0)
    return out, nf, ns, si


# ---- the reference's golden streams (tests/seek.rs:10-31, tests/metadata.rs:30-48) ----
@pytest.mark.parametrize("name,md5", [("sine.flac", "831671b807f97051301e01d68b5c54b3"),
                                      ("all-frames.flac", "f53f86876dcd7783225c93ba8a938c7d"),
                                      ("cuesheet.flac", "2ae74d9f65a6acb8a4e9079125d68952")])
def test_reference_golden_streams(eng, fo, name, md5):
    from flac_codec_b200 import _abi

    flac = ref_file(name)
    out, nf, ns, si = gpu_decode_stream(eng, flac, fo, kind=_abi.PCM_BYTES_LE)
    assert ns == si.total_samples
    assert bytes(si.md5).hex() == md5
    assert hashlib.md5(out[: ns * si.channels * 2].tobytes()).hexdigest() == md5
    if name != "cuesheet.flac":
        ref, _ = fo.decode_stream(flac)
        got = fo.bytes_to_samples(out[: ns * si.channels * 2].tobytes(), 2)
        assert np.array_equal(got, ref)


@pytest.mark.parametrize("chunk", [0, 1, 3, 7])
def test_groups_and_carry(eng, fo, chunk):
    """Small decode groups exercise the chain state that is carried from group to group."""
    flac = ref_file("sine.flac")
    out, nf, ns, si = gpu_decode_stream(eng, flac, fo, chunk=chunk)
    ref, _ = fo.decode_stream(flac)
    assert ns == 200000 and nf == 49
    assert np.array_equal(out.view(np.int32)[: ref.size], ref)
    # packed bytes: the software-pipelined k_parse | CRC + walk + k_restore_emit path, two plane buffers, many groups
    from flac_codec_b200 import _abi

    out, nf, ns, si = gpu_decode_stream(eng, flac, fo, kind=_abi.PCM_BYTES_LE, chunk=chunk)
    assert ns == 200000 and nf == 49
    assert np.array_equal(fo.bytes_to_samples(out[: ref.size * 2].tobytes(), 2), ref)
    # ... and with a damaged frame in the middle: the groups before it are delivered, the error is the reference's
    bad = bytearray(flac)
    bad[si.frames_start + 60000] ^= 0x10
    code, nf_ok, ns_ok, ref_ok = fo.decode_stream_ex(bytes(bad))
    with pytest.raises(_abi.FlacB200Error) as e:
        gpu_decode_stream(eng, bytes(bad), fo, kind=_abi.PCM_BYTES_LE, chunk=chunk)
    assert (e.value.code, e.value.bad_frame) == (code, nf_ok)


CASES = [
    ("default", 44100, 16, 2, 44100 * 2 + 100), ("best", 48000, 24, 2, 48000 + 77), ("default", 44100, 16, 1, 30000),
    ("fast", 44100, 16, 2, 30000), ("best", 96000, 24, 8, 20000), ("best", 192000, 32, 2, 20000), ("default", 8000, 8, 1, 9000),
    ("best", 44100, 12, 2, 9000), ("default", 22050, 20, 3, 9000), ("default", 37, 13, 1, 9000),
]


@pytest.mark.parametrize("preset,rate,bps,ch,n", CASES)
def test_oracle_encoded_streams(eng, fo, preset, rate, bps, ch, n):
    """encode with the oracle (== the reference encoder's bytes), decode on the GPU, compare with the source PCM
    and with the oracle's decoder; all four PCM layouts of Frame::to_buf."""
    from flac_codec_b200 import _abi

    x = synth_pcm(hash((preset, rate, bps, ch)) % 50, ch, n, rate if rate > 1000 else 44100, bps)
    opt = fo.options(preset, max_lpc_order=32) if bps == 32 else fo.options(preset)
    flac, sizes = fo.encode_stream(opt, rate, bps, ch, x.reshape(-1))
    out, nf, ns, si = gpu_decode_stream(eng, flac, fo)
    assert ns == n and nf == len(sizes)
    assert np.array_equal(out.view(np.int32).reshape(-1, ch), x)
    bytes_per = (bps + 7) // 8
    for kind, be in ((_abi.PCM_BYTES_LE, False), (_abi.PCM_BYTES_BE, True)):
        out, _, _, _ = gpu_decode_stream(eng, flac, fo, kind=kind)
        assert out[: n * ch * bytes_per].tobytes() == fo.samples_to_bytes(x.reshape(-1), bytes_per, be)
    out, _, _, _ = gpu_decode_stream(eng, flac, fo, kind=_abi.PCM_I32_PLANAR)
    assert np.array_equal(out.view(np.int32).reshape(ch, -1)[:, :n], x.T)


def test_block_sizes_orders_and_noise(eng, fo):
    """tests/format.rs:85 (block 16..33 x LPC order), noise (escape/verbatim paths), large and odd blocks."""
    data = fo.bytes_to_samples(ref_file("noise32.raw"), 1)
    for blocksize in range(16, 34):
        for lpc_order in [0, 1, 4, 8, 15, 16, 17, 31, 32]:
            opt = fo.options("best", max_lpc_order=lpc_order or None, block_size=blocksize, padding=None)
            flac, _ = fo.encode_stream(opt, 44100, 8, 1, data)
            out, nf, ns, si = gpu_decode_stream(eng, flac, fo)
            assert ns == 32 and np.array_equal(out.view(np.int32)[:32], data), (blocksize, lpc_order)
    rng = np.random.default_rng(5)
    for bps, ch, bs in ((16, 2, 4096), (24, 2, 65535), (32, 1, 32768), (8, 8, 32), (16, 4, 4608)):
        lo, hi = -(1 << (bps - 1)), (1 << (bps - 1)) - 1
        n = 70000 if bs >= 4096 else 1000
        x = rng.integers(lo, hi, size=n * ch, endpoint=True).astype(np.int64).astype(np.int32)
        flac, sizes = fo.encode_stream(fo.options("default", block_size=bs), 44100, bps, ch, x)
        out, nf, ns, si = gpu_decode_stream(eng, flac, fo)
        assert ns == n and nf == len(sizes) and np.array_equal(out.view(np.int32)[: x.size], x), (bps, ch, bs)


def test_wasted_bits_and_full_scale(eng, fo):
    x = fo.bytes_to_samples(ref_file("wasted-bits.raw"), 2)
    flac, _ = fo.encode_stream(fo.options("default"), 44100, 16, 1, x)
    out, nf, ns, si = gpu_decode_stream(eng, flac, fo)
    assert np.array_equal(out.view(np.int32)[: x.size], x)
    for bps in (8, 16, 24, 32):
        hi, lo = (1 << (bps - 1)) - 1, -(1 << (bps - 1))
        for pat in ([hi, lo], [lo, lo, hi], [hi, lo, hi, hi, lo, lo, hi]):
            x = np.array((pat * 1200)[: 4096 + 38], dtype=np.int32)
            for ch in (1, 2):
                flac, _ = fo.encode_stream(fo.options("best"), 44100, bps, ch, x)
                out, nf, ns, si = gpu_decode_stream(eng, flac, fo)
                assert np.array_equal(out.view(np.int32)[: x.size], x), (bps, pat, ch)


def test_multi_segment_batch(eng, fo):
    """Several streams in one call, each with its own byte range and PCM destination (the C4 decode shape)."""
    from flac_codec_b200 import _abi

    rate, bps, ch = 48000, 24, 2
    tracks = [synth_pcm(20 + t, ch, 30000 + 1111 * t, rate, bps) for t in range(6)]
    blobs = [fo.encode_frames_only(fo.options("best"), rate, bps, ch, x.reshape(-1))[0] for x in tracks]
    buf = np.frombuffer(b"".join(blobs), dtype=np.uint8).copy()
    segs, boff, poff = [], 0, 0
    for x, b in zip(tracks, blobs):
        segs.append((boff, len(b), poff, x.shape[0]))
        boff += len(b)
        poff += x.shape[0]
    out = np.zeros(poff * ch, dtype=np.int32)
    for chunk in (0, 5):
        out[:] = 0
        eng.set_chunk_frames(chunk)
        nf, ns = eng.decode(rate, bps, ch, 4096, buf, buf.size, segs, out, out.nbytes, _abi.PCM_I32_INTERLEAVED)
        eng.set_chunk_frames(0)
        assert ns == poff
        assert np.array_equal(out.reshape(-1, ch), np.concatenate(tracks))
    # device-resident input and output
    d_in, d_out = eng.device_alloc(buf.size), eng.device_alloc(out.nbytes)
    eng.memcpy(d_in, buf, buf.size, 1)
    nf, ns = eng.decode(rate, bps, ch, 4096, d_in, buf.size, segs, d_out, out.nbytes, _abi.PCM_I32_INTERLEAVED,
                        frames_location=_abi.DEVICE, pcm_location=_abi.DEVICE)
    out2 = np.zeros_like(out)
    eng.memcpy(out2, d_out, out.nbytes, 2)
    assert np.array_equal(out2, out)
    eng.device_free(d_in)
    eng.device_free(d_out)


def test_batched_host_decode_equals_one_call(eng, fo):
    """Host frames -> host PCM of many streams is cut into batches of segments whose upload, kernels and download overlap
    (decode_batched in engine.cu; 192 MB batches by default, tiny ones here): same PCM, counts and error as one call."""
    from flac_codec_b200 import _abi

    rate, bps, ch = 44100, 16, 2
    tracks = [synth_pcm(40 + t, ch, 9000 + 1234 * t, rate, bps) for t in range(9)]
    blobs = [fo.encode_frames_only(fo.options("default"), rate, bps, ch, x.reshape(-1))[0] for x in tracks]

    def run(buf, segs, total):
        out = np.zeros(total * ch, dtype=np.int32)
        try:
            nf, ns = eng.decode(rate, bps, ch, 4096, buf, buf.size, segs, out, out.nbytes, _abi.PCM_I32_INTERLEAVED)
            return out, nf, ns, 0, 0
        except _abi.FlacB200Error as e:
            return out, None, None, e.code, e.bad_frame

    for damaged in (False, True):
        parts = [bytearray(b) for b in blobs]
        if damaged:
            parts[4][len(parts[4]) // 2] ^= 0x10
        segs, boff, poff = [], 3, 0   # (a gap in front: batch starts are not 16-byte aligned)
        for x, b in zip(tracks, parts):
            segs.append((boff, len(b), poff, x.shape[0]))
            boff += len(b) + 5
            poff += x.shape[0]
        buf = np.zeros(boff, dtype=np.uint8)
        for (o, n, _, _), b in zip(segs, parts):
            buf[o:o + n] = np.frombuffer(bytes(b), dtype=np.uint8)
        eng.set_option("no_batch", 1)
        ref = run(buf, segs, poff)
        eng.set_option("no_batch", 0)
        eng.set_option("batch_bytes", 30000)
        got = run(buf, segs, poff)
        eng.set_option("batch_bytes", 0)
        assert got[1:] == ref[1:], (got[1:], ref[1:])
        if not damaged:
            assert ref[3] == 0 and np.array_equal(got[0], ref[0])
            assert np.array_equal(got[0].reshape(-1, ch), np.concatenate(tracks))
        else:
            assert ref[3] != 0
            good = sum(x.shape[0] for x in tracks[:4]) * ch   # the streams in front of the damaged one
            assert np.array_equal(got[0][:good], ref[0][:good])


# ---- hand-built frames ----
class Bits:
    def __init__(self):
        self.v, self.n = 0, 0

    def put(self, nbits, value):
        self.v = (self.v << nbits) | (value & ((1 << nbits) - 1))
        self.n += nbits

    def bytes(self):
        pad = (-self.n) % 8
        return ((self.v << pad).to_bytes((self.n + pad) // 8, "big"))


def build_verbatim_frame(fo, frame_number, rate_code, bps_code, assignment, block, chans):
    """chans: list of (bits_per_sample, samples) written as VERBATIM subframes."""
    b = Bits()
    b.put(15, 0b111111111111100)
    b.put(1, 0)
    b.put(4, 0b0110 if block <= 256 else 0b0111)
    b.put(4, rate_code)
    b.put(4, assignment)
    b.put(3, bps_code)
    b.put(1, 0)
    assert frame_number < 128
    b.put(8, frame_number)
    b.put(8 if block <= 256 else 16, block - 1)
    hdr = b.bytes()
    b.put(8, fo.crc8(hdr))
    for bits, samples in chans:
        b.put(8, 0b00000010)   # pad, VERBATIM, no wasted bits
        for s in samples:
            b.put(bits, int(s))
    body = b.bytes()
    return body + fo.crc16(body).to_bytes(2, "big")


def test_33_bit_side_channel(eng, fo):
    """32-bit streams with a side channel need 33-bit samples (src/decode.rs:1528-1546, :1565-1583, :1604-1622)."""
    from flac_codec_b200 import _abi

    rng = np.random.default_rng(3)
    n = 200
    left = rng.integers(-(1 << 31), (1 << 31) - 1, size=n, endpoint=True)
    right = rng.integers(-(1 << 31), (1 << 31) - 1, size=n, endpoint=True)
    side = left - right
    mid = (left + right) >> 1
    frames = [build_verbatim_frame(fo, 0, 0b1001, 0b111, 8, n, [(32, left), (33, side)]),
              build_verbatim_frame(fo, 1, 0b1001, 0b111, 9, n, [(33, side), (32, right)]),
              build_verbatim_frame(fo, 2, 0b1001, 0b111, 10, n, [(32, mid), (33, side)])]
    buf = np.frombuffer(b"".join(frames), dtype=np.uint8).copy()
    out = np.zeros(3 * n * 2, dtype=np.int32)
    nf, ns = eng.decode(44100, 32, 2, n, buf, buf.size, [(0, buf.size, 0, 3 * n)], out, out.nbytes, _abi.PCM_I32_INTERLEAVED)
    assert nf == 3 and ns == 3 * n
    want = np.stack([left, right], axis=1).astype(np.int32)
    for k in range(3):
        assert np.array_equal(out.reshape(-1, 2)[k * n:(k + 1) * n], want), k
        planar, h, used = fo.decode_frame(frames[k], None, 0)
        assert np.array_equal(planar.T, want)


def put_residuals(b, rng, nres_first, nparts, chunk, method, kinds):
    """Residual block (src/decode.rs:1800-1856) with one partition kind per partition, cycling through `kinds`:
    ("rice", k, max_quotient) | ("raw", width) | ("zero",)."""
    b.put(2, method)
    porder = nparts.bit_length() - 1
    b.put(4, porder)
    pbits, esc = (5, 31) if method else (4, 15)
    for p in range(nparts):
        cnt = nres_first if p == 0 else chunk
        kind = kinds[p % len(kinds)]
        if kind[0] == "rice":
            k, qmax = kind[1], kind[2]
            b.put(pbits, k)
            for _ in range(cnt):
                q = int(rng.integers(0, qmax + 1))
                b.put(q, 0)
                b.put(1, 1)
                if k:
                    b.put(k, int(rng.integers(0, 1 << k)))
        elif kind[0] == "raw":
            w = kind[1]
            b.put(pbits, esc)
            b.put(5, w)
            for _ in range(cnt):
                b.put(w, int(rng.integers(0, 1 << w)))
        else:
            b.put(pbits, esc)
            b.put(5, 0)


def build_predictive_frame(fo, rng, frame_number, block, order, lpc, porder, method, kinds, wasted=0):
    """One mono 32-bit frame with a FIXED (lpc=None) or LPC (lpc=(precision, shift)) subframe built bit by bit."""
    b = Bits()
    b.put(15, 0b111111111111100)
    b.put(1, 0)
    b.put(4, 0b0110 if block <= 256 else 0b0111)
    b.put(4, 0b1001)
    b.put(4, 0)        # one channel
    b.put(3, 0b111)    # 32 bits per sample
    b.put(1, 0)
    b.put(8, frame_number)
    b.put(8 if block <= 256 else 16, block - 1)
    b.put(8, fo.crc8(b.bytes()))
    b.put(1, 0)
    b.put(6, (31 + order) if lpc else (8 + order))
    if wasted:
        b.put(1, 1)
        b.put(wasted - 1, 0)
        b.put(1, 1)
    else:
        b.put(1, 0)
    ebps = 32 - wasted
    for _ in range(order):
        b.put(ebps, int(rng.integers(0, 1 << min(ebps, 20))))
    if lpc:
        prec, shift = lpc
        b.put(4, prec - 1)
        b.put(5, shift)
        for _ in range(order):
            b.put(prec, int(rng.integers(0, 1 << prec)))
    nparts, chunk = 1 << porder, block >> porder
    put_residuals(b, rng, chunk - order, nparts, chunk, method, kinds)
    body = b.bytes()
    return body + fo.crc16(body).to_bytes(2, "big")


def test_hand_built_predictive_frames(eng, fo):
    """Shapes no encoder here emits: one-sample partitions, Rice parameters up to 30, unary runs longer than the bit window,
    raw (escaped) partitions of every width, LPC order 32 with 15-bit coefficients, wasted bits -- the GPU must produce
    what the oracle's serial decoder produces for the same bytes."""
    from flac_codec_b200 import _abi

    rng = np.random.default_rng(99)
    all_rice = [("rice", k, 3) for k in range(0, 31)]
    long_runs = [("rice", 0, 200), ("rice", 3, 90), ("rice", 14, 40), ("rice", 30, 5)]
    raws = [("raw", w) for w in range(1, 32)] + [("zero",)]
    frames = [
        build_predictive_frame(fo, rng, 0, 256, 0, None, 8, 1, all_rice + raws),               # one sample per partition
        build_predictive_frame(fo, rng, 1, 256, 1, None, 7, 1, raws + long_runs),              # first partition: one residual
        build_predictive_frame(fo, rng, 2, 256, 4, None, 5, 0, [("rice", k, 6) for k in range(15)] + [("raw", 9), ("zero",)]),
        build_predictive_frame(fo, rng, 3, 256, 32, (15, 14), 2, 1, long_runs + [("raw", 31)]),
        build_predictive_frame(fo, rng, 4, 256, 32, (15, 0), 0, 0, [("rice", 7, 70)]),
        build_predictive_frame(fo, rng, 5, 256, 12, (12, 9), 3, 1, all_rice, wasted=5),
        build_predictive_frame(fo, rng, 6, 256, 2, None, 6, 1, [("raw", 31), ("raw", 1), ("rice", 29, 2)], wasted=1),
    ]
    want = []
    for f in frames:
        planar, h, used = fo.decode_frame(f, None, 0)
        assert used == len(f)
        want.append(planar[0])
    buf = np.frombuffer(b"".join(frames), dtype=np.uint8).copy()
    n = 256 * len(frames)
    out = np.zeros(n, dtype=np.int32)
    nf, ns = eng.decode(44100, 32, 1, 256, buf, buf.size, [(0, buf.size, 0, n)], out, out.nbytes, _abi.PCM_I32_INTERLEAVED)
    assert (nf, ns) == (len(frames), n)
    for k, w in enumerate(want):
        assert np.array_equal(out[k * 256:(k + 1) * 256], w), k
    # the same frames many times over, so that whole warps walk them in lockstep at different phases
    reps = 70
    order = rng.integers(0, len(frames), size=reps)
    rebuilt = []   # re-numbered 0..69: CRC-8 and CRC-16 have to be rebuilt
    for k, i in enumerate(order):
        f = bytearray(frames[int(i)])
        f[4] = k
        f[6] = fo.crc8(bytes(f[:6]))
        f[-2:] = fo.crc16(bytes(f[:-2])).to_bytes(2, "big")
        rebuilt.append(bytes(f))
    buf = np.frombuffer(b"".join(rebuilt), dtype=np.uint8).copy()
    out = np.zeros(256 * reps, dtype=np.int32)
    nf, ns = eng.decode(44100, 32, 1, 256, buf, buf.size, [(0, buf.size, 0, 256 * reps)], out, out.nbytes, _abi.PCM_I32_INTERLEAVED)
    assert (nf, ns) == (reps, 256 * reps)
    for k, i in enumerate(order):
        assert np.array_equal(out[k * 256:(k + 1) * 256], want[int(i)]), (k, int(i))


def test_sync_code_inside_payload_is_skipped(eng, fo):
    """A complete, CRC-8-valid frame header embedded in a VERBATIM payload is a false candidate: the walk that
    follows frame ends must skip it, exactly as the reference's serial reader never sees it."""
    from flac_codec_b200 import _abi

    n = 64
    inner = build_verbatim_frame(fo, 7, 0b1001, 0b100, 0, n, [(16, np.arange(n))])
    fake_hdr = inner[:7]   # ff f8 .. crc8 : 7 bytes; pad to whole 16-bit samples
    payload = list(np.arange(10, 10 + n))
    raw = fake_hdr + b"\x00"
    words = [int.from_bytes(raw[i:i + 2], "big", signed=True) for i in range(0, 8, 2)]
    payload[8:12] = words
    f0 = build_verbatim_frame(fo, 0, 0b1001, 0b100, 0, n, [(16, payload)])
    f1 = build_verbatim_frame(fo, 1, 0b1001, 0b100, 0, n, [(16, np.arange(n) * 3)])
    assert fake_hdr in f0
    buf = np.frombuffer(f0 + f1, dtype=np.uint8).copy()
    out = np.zeros(2 * n, dtype=np.int32)
    nf, ns = eng.decode(44100, 16, 1, n, buf, buf.size, [(0, buf.size, 0, 2 * n)], out, out.nbytes, _abi.PCM_I32_INTERLEAVED)
    assert nf == 2 and ns == 2 * n
    assert out[:n].tolist() == payload and out[n:].tolist() == (np.arange(n) * 3).tolist()


# ---- tests/corruption.rs: one flipped bit must be detected ----
def _corrupt_streams(fo):
    yield "sine.flac", bytearray(ref_file("sine.flac")), 1000
    for name in ("all-frames.flac", "cuesheet.flac", "seektable.flac"):
        yield name, bytearray(ref_file(name)), 150
    for bps, ch, preset, n, kw in ((24, 2, "best", 40000, {}), (32, 2, "best", 12000, {}), (24, 8, "best", 9000, {}),
                                   (16, 1, "fast", 20000, {}), (8, 2, "default", 6000, {"block_size": 64})):
        x = synth_pcm(11, ch, n, 48000, bps)
        flac, _ = fo.encode_stream(fo.options(preset, **kw), 48000, bps, ch, x.reshape(-1))
        yield f"synth {bps}b {ch}ch {preset}", bytearray(flac), 150


def test_bit_flips_give_the_reference_verdict(eng, fo):
    """Single-bit corruptions anywhere in the frame data (headers included): the GPU must return the SAME Error ordinal
    and the SAME failing frame index as the serial reader (Decoder::read_frame, src/decode.rs:1388; header checks in the
    order of src/stream.rs:214-313), and deliver the same samples in front of the failing frame -- in every trial
    (tools/corrupt_probe.py: 4600 flips over these streams, both decoders, no disagreement)."""
    from flac_codec_b200 import _abi

    rng = np.random.default_rng(1234)
    for name, flac, trials in _corrupt_streams(fo):
        si = fo.read_streaminfo(bytes(flac))
        for t in range(trials):
            pos = int(rng.integers(si.frames_start * 8, len(flac) * 8))
            flac[pos >> 3] ^= 0x80 >> (pos & 7)
            code, nf, ns, ref = fo.decode_stream_ex(bytes(flac))
            frames = np.frombuffer(bytes(flac), dtype=np.uint8)[si.frames_start:].copy()
            out = np.zeros(si.total_samples * si.channels, dtype=np.int32)
            got = (0, nf)
            try:
                eng.decode(si.sample_rate, si.bps, si.channels, si.max_block_size, frames, frames.size,
                           [(0, frames.size, 0, si.total_samples)], out, out.nbytes, _abi.PCM_I32_INTERLEAVED)
            except _abi.FlacB200Error as e:
                got = (e.code, e.bad_frame)
            assert got == (code, nf), (name, t, pos, got, (code, nf))
            assert np.array_equal(out[:ns], ref), (name, t, pos)
            if si.bps <= 24 and si.channels <= 2 and t % 2 == 0:   # the packed-bytes layout (k_restore_emit after whichever walk ran)
                nb = (si.bps + 7) // 8
                outb = np.zeros(si.total_samples * si.channels * nb, dtype=np.uint8)
                got = (0, nf)
                try:
                    eng.decode(si.sample_rate, si.bps, si.channels, si.max_block_size, frames, frames.size,
                               [(0, frames.size, 0, si.total_samples)], outb, outb.nbytes, _abi.PCM_BYTES_LE)
                except _abi.FlacB200Error as e:
                    got = (e.code, e.bad_frame)
                assert got == (code, nf), (name, t, pos, got, (code, nf), "bytes")
                assert np.array_equal(fo.bytes_to_samples(outb[: ns * nb].tobytes(), nb), ref), (name, t, pos, "bytes")
            flac[pos >> 3] ^= 0x80 >> (pos & 7)


def test_stream_end_rules(eng, fo):
    """ShortBlock (src/decode.rs:1405-1410), early end with a known total (Io), unknown total reads to the end."""
    from flac_codec_b200 import _abi

    x = synth_pcm(30, 1, 4096 * 3 + 10, 44100, 16)
    frames, sizes = fo.encode_frames_only(fo.options("default"), 44100, 16, 1, x.reshape(-1))
    buf = np.frombuffer(frames, dtype=np.uint8).copy()
    n = x.shape[0]
    out = np.zeros(n + 5000, dtype=np.int32)
    # total unknown: decode until the bytes end
    nf, ns = eng.decode(44100, 16, 1, 4096, buf, buf.size, [(0, buf.size, 0, 0)], out, out.nbytes, _abi.PCM_I32_INTERLEAVED)
    assert (nf, ns) == (4, n) and np.array_equal(out[:n], x[:, 0])
    # the 10-sample last block is legal only as the very last block
    with pytest.raises(_abi.FlacB200Error) as ei:
        eng.decode(44100, 16, 1, 4096, buf, buf.size, [(0, buf.size, 0, n + 1)], out, out.nbytes, _abi.PCM_I32_INTERLEAVED)
    assert ei.value.code == 21 and ei.value.bad_frame == 3
    # known total, bytes end early -> UnexpectedEof
    cut = int(sizes[:2].sum())
    with pytest.raises(_abi.FlacB200Error) as ei:
        eng.decode(44100, 16, 1, 4096, buf, cut, [(0, cut, 0, n)], out, out.nbytes, _abi.PCM_I32_INTERLEAVED)
    assert ei.value.code == 1 and ei.value.bad_frame == 2
    # known total smaller than the stream: extra frames are never read
    nf, ns = eng.decode(44100, 16, 1, 4096, buf, buf.size, [(0, buf.size, 0, 8192)], out, out.nbytes, _abi.PCM_I32_INTERLEAVED)
    assert (nf, ns) == (2, 8192)
    # garbage instead of the first frame
    bad = buf.copy()
    bad[0] = 0
    with pytest.raises(_abi.FlacB200Error) as ei:
        eng.decode(44100, 16, 1, 4096, bad, bad.size, [(0, bad.size, 0, n)], out, out.nbytes, _abi.PCM_I32_INTERLEAVED)
    assert ei.value.code == 23 and ei.value.bad_frame == 0   # InvalidSyncCode
    # STREAMINFO cross-checks (src/stream.rs:291-312)
    with pytest.raises(_abi.FlacB200Error) as ei:
        eng.decode(48000, 16, 1, 4096, buf, buf.size, [(0, buf.size, 0, n)], out, out.nbytes, _abi.PCM_I32_INTERLEAVED)
    assert ei.value.code == 29   # SampleRateMismatch
    with pytest.raises(_abi.FlacB200Error) as ei:
        eng.decode(44100, 16, 1, 1024, buf, buf.size, [(0, buf.size, 0, n)], out, out.nbytes, _abi.PCM_I32_INTERLEAVED)
    assert ei.value.code == 25   # BlockSizeMismatch
    # output too small
    with pytest.raises(_abi.FlacB200Error) as ei:
        eng.decode(44100, 16, 1, 4096, buf, buf.size, [(0, buf.size, 0, n)], out, 4 * 5000, _abi.PCM_I32_INTERLEAVED)
    assert ei.value.code == -4


def test_encode_decode_round_trip_at_bench_shape(eng, fo):
    """The C4 shape end to end on the device: synth -> encode -> decode -> identical PCM bytes (size-independent
    property used at BASELINE.json's full sizes by bench.py)."""
    from flac_codec_b200 import Options, _abi

    rate, bps, ch, n, tracks = 48000, 24, 2, 48000 * 4 + 999, 12
    nbytes = tracks * n * ch * 3
    d_pcm = eng.device_alloc(nbytes)
    eng.synth_pcm(d_pcm, 0, tracks, n, ch, rate, bps)
    d_flac = eng.device_alloc(nbytes)
    segs = [(t * n, n, 0) for t in range(tracks)]
    _, sizes, total = eng.encode(Options.best(), rate, bps, ch, d_pcm, nbytes, _abi.PCM_BYTES_LE, segs, pcm_location=_abi.DEVICE,
                                 out=d_flac, out_capacity=nbytes, out_location=_abi.DEVICE)
    per_track = (n + 4095) // 4096
    offs = np.concatenate([[0], np.cumsum(sizes.astype(np.int64))])
    dsegs = [(int(offs[t * per_track]), int(offs[(t + 1) * per_track] - offs[t * per_track]), t * n, n) for t in range(tracks)]
    d_back = eng.device_alloc(nbytes)
    nf, ns = eng.decode(rate, bps, ch, 4096, d_flac, total, dsegs, d_back, nbytes, _abi.PCM_BYTES_LE, frames_location=_abi.DEVICE,
                        pcm_location=_abi.DEVICE)
    assert nf == tracks * per_track and ns == tracks * n
    a, b = np.zeros(nbytes, dtype=np.uint8), np.zeros(nbytes, dtype=np.uint8)
    eng.memcpy(a, d_pcm, nbytes, 2)
    eng.memcpy(b, d_back, nbytes, 2)
    assert np.array_equal(a, b)
    for p in (d_pcm, d_flac, d_back):
        eng.device_free(p)


def test_last_frame_overshoots_an_understated_total(eng, fo, decode_path):
    """STREAMINFO announces fewer samples than the frames hold: `total - current_sample` (src/decode.rs:1400) passes the
    ShortBlock rule for a block of more than 14 samples, the frame is delivered, the difference wraps (release build) and the
    reader runs into the end of the bytes: every frame comes out, then Io.  The GPU walk and the reader handle do the same."""
    import io

    from flac_codec_b200 import _abi, stream as st

    x = synth_pcm(31, 2, 4096 * 5 + 3000, 44100, 16).reshape(-1)
    flac, sizes = fo.encode_stream(fo.options("default", padding=None), 44100, 16, 2, x, total_known=True)
    si = fo.read_streaminfo(flac)
    for short_by in (100, 4096 * 2 + 50):
        lied = bytearray(flac)
        v = int.from_bytes(lied[4 + 4 + 10:4 + 4 + 18], "big")
        total = (v & 0xFFFFFFFFF) - short_by
        lied[4 + 4 + 10:4 + 4 + 18] = ((v & ~0xFFFFFFFFF) | total).to_bytes(8, "big")
        code, nf, ns, ref = fo.decode_stream_ex(bytes(lied))
        assert (code, nf) == (1, len(sizes)) and np.array_equal(ref, x)      # all frames, then Io
        frames = np.frombuffer(bytes(lied), dtype=np.uint8)[si.frames_start:].copy()
        out = np.zeros(x.size + 8192, dtype=np.int32)
        with pytest.raises(_abi.FlacB200Error) as ei:
            eng.decode(44100, 16, 2, 4096, frames, frames.size, [(0, frames.size, 0, total)], out, out.nbytes, _abi.PCM_I32_INTERLEAVED)
        assert (ei.value.code, ei.value.bad_frame) == (1, len(sizes))
        assert np.array_equal(out[: x.size], x)
        r = st.FlacSampleReader(bytes(lied), engine=eng, window_bytes=20000)
        got = []
        with pytest.raises(_abi.FlacB200Error) as ei:
            while True:
                a = r.read(10000)
                if a.size == 0:
                    break
                got.append(a.copy())
        assert ei.value.code == 1 and np.array_equal(np.concatenate(got), x)
        r.close()


def test_single_pass_frame_discovery_equals_two_passes(eng, fo, decode_path):
    """k_find parks a tile's candidates while it counts them (the bytes are read once); a tile with more candidates than
    slots -- streams of tiny blocks -- sends the call back to the second pass.  Both give the same frames (legacy bit 256 forces
    the two-pass form)."""
    from flac_codec_b200 import _abi

    base = 64 if decode_path == "thread-per-frame" else 0
    cases = [(synth_pcm(50, 2, 4096 * 40 + 99, 44100, 16), fo.options("default"), 44100, 16),              # one or two candidates per tile
             (synth_pcm(51, 1, 16 * 3000 + 5, 44100, 16), fo.options("default", block_size=16), 44100, 16),   # hundreds per tile: second pass
             (synth_pcm(52, 2, 192 * 500, 48000, 24), fo.options("best", block_size=192), 48000, 24)]
    for x, opt, rate, bps in cases:
        ch = x.shape[1]
        frames, sizes = fo.encode_frames_only(opt, rate, bps, ch, x.reshape(-1))
        buf = np.frombuffer(frames, dtype=np.uint8).copy()
        outs = []
        for legacy in (base, base | 256):
            eng.set_option("legacy", legacy)
            out = np.zeros(x.size, dtype=np.int32)
            nf, ns = eng.decode(rate, bps, ch, opt.block_size, buf, buf.size, [(0, buf.size, 0, x.shape[0])], out, out.nbytes, _abi.PCM_I32_INTERLEAVED)
            assert (nf, ns) == (len(sizes), x.shape[0])
            outs.append(out)
        eng.set_option("legacy", base)
        assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0].reshape(-1, ch), x)
