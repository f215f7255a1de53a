"""GPU encode parity (through the C ABI): frames must be byte-identical to the CPU oracle's for the
same blocks and options, and must decode losslessly through the oracle's (independently pinned)
decoder.  The matrix replays the reference's tests/format.rs cases."""
import os
import sys

import numpy as np
import pytest

from flacb200_testutil import ROOT, generate_sine_1, generate_sine_2, ref_file, synth_pcm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from flac_codec_b200 import Engine

    e = Engine(0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def fo():
    from oracle import oracle

    return oracle


def _probe():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import parity_probe

    return parity_probe


def check(eng, fo, opt, rate, bps, channels, x, label="", first_frame_number=0, pcm_kind=None):
    """GPU frames == oracle frames, and oracle-decode(GPU frames) == x."""
    x = np.ascontiguousarray(x, dtype=np.int32).reshape(-1)
    same, n, gpu, ref = _probe().compare(eng, opt, rate, bps, channels, x, label, first_frame_number, verbose=True,
                                         pcm_kind=pcm_kind)
    assert same == n, f"{label}: {same}/{n} frames identical"
    assert gpu == ref
    # lossless through the oracle decoder, frame by frame (subset semantics: no STREAMINFO needed
    # when the header carries rate and bps; otherwise give it a STREAMINFO)
    si = fo.Streaminfo()
    si.min_block_size = si.max_block_size = opt.c.block_size
    si.sample_rate, si.channels, si.bps = rate, channels, bps
    si.total_samples = x.size // channels
    pos, done = 0, 0
    total = x.size // channels
    while done < total:
        planar, h, used = fo.decode_frame(gpu[pos:], si, total - done)
        blk = x.reshape(-1, channels)[done:done + h.block_size].T
        assert np.array_equal(planar, blk), f"{label}: frame at byte {pos} does not round-trip"
        pos += used
        done += h.block_size
    assert pos == len(gpu)


def test_bench_signals_all_presets(eng, fo):
    from flac_codec_b200 import Options

    cases = [
        ("16b stereo default", Options.default(), 44100, 16, 2, synth_pcm(0, 2, 44100 * 2 + 100, 44100, 16)),
        ("24b stereo best", Options.best(), 48000, 24, 2, synth_pcm(1, 2, 48000 * 2 + 77, 48000, 24)),
        ("16b mono default", Options.default(), 44100, 16, 1, synth_pcm(2, 1, 50000, 44100, 16)),
        ("16b stereo fast", Options.fast(), 44100, 16, 2, synth_pcm(3, 2, 50000, 44100, 16)),
        ("24b 8ch best", Options.best(), 96000, 24, 8, synth_pcm(4, 8, 20000, 96000, 24)),
        ("32b stereo best lpc32", Options.best().max_lpc_order(32), 192000, 32, 2, synth_pcm(5, 2, 30000, 192000, 32)),
        ("16b stereo no-mid-side", Options.default().mid_side(False), 44100, 16, 2, synth_pcm(6, 2, 30000, 44100, 16)),
        ("16b stereo fast-corr mid-side", Options.default().fast_channel_correlation(True), 44100, 16, 2,
         synth_pcm(7, 2, 30000, 44100, 16)),
        ("16b stereo hann", Options.default().window("hann"), 44100, 16, 2, synth_pcm(8, 2, 20000, 44100, 16)),
        ("16b stereo rectangle", Options.default().window("rectangle"), 44100, 16, 2, synth_pcm(9, 2, 20000, 44100, 16)),
        ("12b mono streaminfo-bps", Options.default(), 37, 13, 1, synth_pcm(10, 1, 9000, 44100, 13)),
    ]
    for label, opt, rate, bps, ch, x in cases:
        check(eng, fo, opt, rate, bps, ch, x, label)


# tests/format.rs:208 test_roundtrip (36 fixtures x 3 presets)
@pytest.mark.parametrize("channels", [1, 2, 4, 8])
@pytest.mark.parametrize("bps", [8, 16, 24])
def test_reference_roundtrip_fixtures(eng, fo, channels, bps):
    from flac_codec_b200 import Options

    for frames in (1, 111, 4777):
        raw = ref_file(f"roundtrip-{channels}-{bps}-{frames}.raw")
        x = fo.bytes_to_samples(raw, bps // 8)
        for preset in ("default", "fast", "best"):
            check(eng, fo, Options(preset), 44100, bps, channels, x, f"roundtrip-{channels}-{bps}-{frames} {preset}")


# tests/format.rs:85 test_blocksize_variations
def test_blocksize_variations(eng, fo):
    from flac_codec_b200 import Options

    data = fo.bytes_to_samples(ref_file("noise32.raw"), 1)
    for blocksize in range(16, 34):
        for lpc_order in [0, 1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 31, 32]:
            opt = Options.best().max_lpc_order(lpc_order or None).block_size(blocksize)
            check(eng, fo, opt, 44100, 8, 1, data, f"bs{blocksize} lpc{lpc_order}")


# tests/format.rs:137 test_fractional
def test_fractional(eng, fo):
    from flac_codec_b200 import Options

    rng = np.random.default_rng(42)
    noise = rng.integers(-32768, 32767, size=16390 * 2, endpoint=True).astype(np.int32)
    cases = [(33, [31, 32, 33, 34, 35, 2046, 2049]), (256, [254, 255, 256, 257, 258, 511, 513, 4098]),
             (2048, [1022, 2047, 2048, 2049, 4097]), (4608, [1023, 4607, 4608, 4609, 8193, 16386])]
    for blocksize, counts in cases:
        opt = Options.default().block_size(blocksize)
        for samples in counts:
            check(eng, fo, opt, 44100, 16, 2, noise[: samples * 2], f"fractional bs{blocksize} n{samples}")


# tests/format.rs:438 test_full_scale_deflection
@pytest.mark.parametrize("bps", [8, 16, 24, 32])
def test_full_scale_deflection(eng, fo, bps):
    from flac_codec_b200 import Options

    hi, lo = (1 << (bps - 1)) - 1, -(1 << (bps - 1))
    patterns = [[hi] * 2, [lo] * 2, [hi, lo], [lo, hi], [hi, hi, lo], [lo, lo, hi], [hi, lo, lo], [lo, hi, hi],
                [hi, hi, lo, lo], [hi, lo, hi, hi, lo, lo, hi]]
    for k, pat in enumerate(patterns):
        x = np.array((pat * 1200)[:4096 + 37], dtype=np.int32)
        for preset in ("default", "best"):
            check(eng, fo, Options(preset), 44100, bps, 1, x, f"fsd{bps} p{k} mono {preset}")
            check(eng, fo, Options(preset), 44100, bps, 2, x[: 2 * (len(x) // 2)], f"fsd{bps} p{k} stereo {preset}")


# tests/format.rs:624 test_wasted_bits
def test_wasted_bits(eng, fo):
    from flac_codec_b200 import Options

    x = fo.bytes_to_samples(ref_file("wasted-bits.raw"), 2)
    check(eng, fo, Options.default(), 44100, 16, 1, x, "wasted-bits")
    infos, n = eng.last_info()
    assert n == 1 and infos[0].sub[0].wasted == 2


# tests/format.rs:777 test_sine_wave_streams
@pytest.mark.parametrize("bps", [8, 16, 24, 32])
def test_sine_streams(eng, fo, bps):
    from flac_codec_b200 import Options

    fs = float((1 << (bps - 1)) - 1)
    for f1, a1, f2, a2 in [(441.0, 0.50, 441.0, 0.49), (441.0, 0.61, 661.5, 0.37), (8820.0, 0.70, 4410.0, 0.29)]:
        x = generate_sine_1(fs, 48000.0, 20000, f1, a1, f2, a2)
        check(eng, fo, Options.default(), 48000, bps, 1, x, f"sine1 {bps} {f1}")
    for f1, a1, f2, a2, fm in [(441.0, 0.50, 441.0, 0.49, 1.0), (441.0, 0.61, 661.5, 0.37, 2.0),
                               (8820.0, 0.70, 4410.0, 0.29, 0.5)]:
        x = generate_sine_2(fs, 44100.0, 20000, f1, a1, f2, a2, fm)
        for preset in ("default", "best", "fast"):
            check(eng, fo, Options(preset), 44100, bps, 2, x, f"sine2 {bps} {f1} {preset}")


# tests/format.rs:1248-1384 test_noise_*
@pytest.mark.parametrize("bps", [8, 16, 24, 32])
@pytest.mark.parametrize("channels", [1, 2, 4, 8])
def test_noise(eng, fo, bps, channels):
    from flac_codec_b200 import Options

    rng = np.random.default_rng(bps * 10 + channels)
    lo, hi = -(1 << (bps - 1)), (1 << (bps - 1)) - 1
    n = 70000
    x = rng.integers(lo, hi, size=n * channels, endpoint=True).astype(np.int64).astype(np.int32)
    for preset, bs in (("default", 4096), ("fast", 32), ("best", 32768), ("default", 65535)):
        m = n if bs >= 4096 else 1000
        check(eng, fo, Options(preset).block_size(bs), 44100, bps, channels, x[: m * channels], f"noise {bps}/{channels} {preset} {bs}")


def test_pcm_layouts_and_frame_numbers(eng, fo):
    """bytes LE/BE, i32 interleaved and planar inputs give the same frames; frame numbers beyond one
    UTF-8 byte are coded like the reference (src/stream.rs:1266-1326)."""
    from flac_codec_b200 import Options, _abi

    x = synth_pcm(11, 2, 30000, 48000, 24)
    for kind in (_abi.PCM_BYTES_LE, _abi.PCM_BYTES_BE, _abi.PCM_I32_INTERLEAVED):
        check(eng, fo, Options.best(), 48000, 24, 2, x, f"layout {kind}", pcm_kind=kind)
    check(eng, fo, Options.default(), 44100, 16, 2, synth_pcm(12, 2, 9000, 44100, 16), "fn 127", first_frame_number=126)
    check(eng, fo, Options.default(), 44100, 16, 2, synth_pcm(12, 2, 9000, 44100, 16), "fn 2^31", first_frame_number=(1 << 31) - 1)
    # planar
    planar = np.ascontiguousarray(x.T)
    ref, ref_sizes = fo.encode_frames_only(fo.options("best"), 48000, 24, 2, x.reshape(-1))
    data, sizes, total = eng.encode(Options.best(), 48000, 24, 2, planar, planar.nbytes, _abi.PCM_I32_PLANAR,
                                    [(0, x.shape[0], 0)], planar_stride=x.shape[0])
    assert data.tobytes() == ref


def test_multi_segment_batch_and_device_buffers(eng, fo):
    """Several tracks in one call (the C4 shape), PCM generated on the device, output left on the device."""
    from flac_codec_b200 import Options, _abi

    rate, bps, ch, n, tracks = 48000, 24, 2, 48000 + 1234, 5
    nbytes = tracks * n * ch * 3
    d_pcm = eng.device_alloc(nbytes)
    eng.synth_pcm(d_pcm, 3, tracks, n, ch, rate, bps)
    host = np.zeros(nbytes, dtype=np.uint8)
    eng.memcpy(host, d_pcm, nbytes, 2)
    # the device generator is bit-identical to the numpy statement
    for t in range(tracks):
        want = synth_pcm(3 + t, ch, n, rate, bps)
        got = fo.bytes_to_samples(host[t * n * ch * 3:(t + 1) * n * ch * 3].tobytes(), 3).reshape(-1, ch)
        assert np.array_equal(got, want), f"synth track {t}"
    segs = [(t * n, n, 0) for t in range(tracks)]
    opt = Options.best()
    cap = 2 * nbytes
    d_out = eng.device_alloc(cap)
    _, sizes, total = eng.encode(opt, rate, bps, ch, d_pcm, nbytes, _abi.PCM_BYTES_LE, segs, pcm_location=_abi.DEVICE,
                                 out=d_out, out_capacity=cap, out_location=_abi.DEVICE)
    got = np.zeros(total, dtype=np.uint8)
    eng.memcpy(got, d_out, total, 2)
    ref = b""
    for t in range(tracks):
        r, _ = fo.encode_frames_only(fo.options("best"), rate, bps, ch, synth_pcm(3 + t, ch, n, rate, bps).reshape(-1))
        ref += r
    assert got.tobytes() == ref
    eng.device_free(d_pcm)
    eng.device_free(d_out)
    # small launch groups give the same bytes
    eng.set_chunk_frames(7)
    data, sizes2, total2 = eng.encode(opt, rate, bps, ch, host, nbytes, _abi.PCM_BYTES_LE, segs)
    eng.set_chunk_frames(0)
    assert data.tobytes() == ref and sizes2.tolist() == sizes.tolist()


@pytest.mark.parametrize("aligned", [True, False])
def test_device_pcm_ragged_segments_match_host_pcm(eng, fo, aligned):
    """Device-resident PCM: the frame table is built on the device (k_descs) and the LPC analysis of the whole call is one
    launch (k_lpc4).  Ragged segments -- shorter than a block, a tail block, empty, out of order in the buffer, frame
    numbers that do not start at 0 -- must give the bytes of the host-PCM call (host-built table, per-group launches)
    and of the oracle, for one launch group and for many."""
    from flac_codec_b200 import Options, _abi

    rate, bps, ch = 44100, 16, 2
    bs = Options.best().c.block_size
    # aligned: every block starts on a 16-byte boundary (k_lpc4's cp.async staging); otherwise the unstaged LPC kernel runs
    lens = [3 * bs + 16, 4, bs, 0, 2 * bs, bs - 4, 16, 7 * bs + 4000] if aligned else [3 * bs + 17, 5, bs, 0, 2 * bs, bs - 1, 16, 7 * bs + 4000]
    total = sum(lens) + 64
    x = synth_pcm(11, ch, total, rate, bps)
    raw = np.frombuffer(fo.samples_to_bytes(x.reshape(-1), 2), dtype=np.uint8).copy()
    offs, pos = [], 32
    for n in lens:
        offs.append(pos)
        pos += n
    order = [4, 0, 7, 3, 1, 6, 2, 5]                 # segments in an order that is not the buffer's
    firsts = [0, 5, 1000, 7, 2 ** 31 - 2, 0, 3, 9]   # first frame number of each
    segs = [(offs[i], lens[i], firsts[i]) for i in order]
    ref = b""
    for i in order:
        if lens[i]:
            r, _ = fo.encode_frames_only(fo.options("best"), rate, bps, ch, x[offs[i]:offs[i] + lens[i]].reshape(-1), first_frame_number=firsts[i])
            ref += r
    d_pcm = eng.device_alloc(raw.nbytes)
    eng.memcpy(d_pcm, raw, raw.nbytes, 1)
    for chunk in (0, 3):
        eng.set_chunk_frames(chunk)
        dev, sizes_d, total_d = eng.encode(Options.best(), rate, bps, ch, d_pcm, raw.nbytes, _abi.PCM_BYTES_LE, segs, pcm_location=_abi.DEVICE)
        host, sizes_h, total_h = eng.encode(Options.best(), rate, bps, ch, raw, raw.nbytes, _abi.PCM_BYTES_LE, segs)
        eng.set_chunk_frames(0)
        assert dev.tobytes() == ref, f"device PCM, chunk {chunk}"
        assert host.tobytes() == ref, f"host PCM, chunk {chunk}"
        assert sizes_d.tolist() == sizes_h.tolist() and total_d == total_h == len(ref)
    eng.device_free(d_pcm)


def test_lpc_overlap_option_same_bytes(eng, fo):
    """Option lpc_overlap: the LPC analysis of launch group g + 1 as a persistent grid (k_lpc4's two-warp form, or k_lpc3)
    on a second stream beside the integer kernels of group g -- off by default, byte-identical."""
    from flac_codec_b200 import Options

    x = synth_pcm(21, 2, 4096 * 37 + 500, 48000, 24)
    for legacy in (0, 2048):
        eng.set_option("legacy", legacy)
        eng.set_option("lpc_overlap", 1)
        eng.set_chunk_frames(5)
        try:
            check(eng, fo, Options.best(), 48000, 24, 2, x, f"lpc_overlap legacy={legacy}")
        finally:
            eng.set_chunk_frames(0)
            eng.set_option("lpc_overlap", 0)
            eng.set_option("legacy", 0)


def test_subset_stream_writer_semantics(eng, fo):
    """FlacStreamWriter::write (src/encode.rs:1094): subset header rules and errors."""
    from flac_codec_b200 import Options, _abi

    x = synth_pcm(13, 2, 4096, 44100, 16)
    raw = np.frombuffer(fo.samples_to_bytes(x.reshape(-1), 2), dtype=np.uint8).copy()
    data, sizes, total = eng.encode(Options.default(), 44100, 16, 2, raw, raw.nbytes, _abi.PCM_BYTES_LE, [(0, 4096, 5)],
                                    subset=True)
    planar = np.ascontiguousarray(x.T)
    ref = fo.encode_frame(fo.options("default"), 44100, 16, planar, frame_number=5, subset=True)
    assert data.tobytes() == ref
    with pytest.raises(_abi.FlacB200Error) as ei:
        eng.encode(Options.default(), 44100, 13, 2, raw, raw.nbytes, _abi.PCM_BYTES_LE, [(0, 4096, 0)], subset=True)
    assert ei.value.code == 28   # NonSubsetBitsPerSample


def test_argument_errors(eng):
    from flac_codec_b200 import Options, _abi

    raw = np.zeros(64, dtype=np.uint8)
    with pytest.raises(_abi.FlacB200Error) as ei:
        eng.encode(Options.default(), 44100, 16, 9, raw, raw.nbytes, _abi.PCM_BYTES_LE, [(0, 1, 0)])
    assert ei.value.code == 30   # ExcessiveChannels
    with pytest.raises(_abi.FlacB200Error) as ei:
        eng.encode(Options.default(), 44100, 33, 2, raw, raw.nbytes, _abi.PCM_BYTES_LE, [(0, 1, 0)])
    assert ei.value.code == 33   # InvalidBitsPerSample
    with pytest.raises(_abi.FlacB200Error) as ei:
        eng.encode(Options.default(), 44100, 16, 2, raw, raw.nbytes, _abi.PCM_BYTES_LE, [(0, 1000, 0)])
    assert ei.value.code == -2   # pcm buffer too small
    # empty input: no frames, no error
    data, sizes, total = eng.encode(Options.default(), 44100, 16, 2, raw, 0, _abi.PCM_BYTES_LE, [(0, 0, 0)])
    assert total == 0 and len(sizes) == 0


def test_generic_kernels_give_the_same_bytes(eng, fo):
    """The engine picks register-tiled kernels for blocks <= 4096 / samples <= 28 bits / LPC order <= 16 and generic
    ones otherwise; FLACB200_LEGACY forces the generic kernels so that both paths are checked on the same inputs."""
    from flac_codec_b200 import Options

    cases = [
        ("16b stereo default", Options.default(), 44100, 16, 2, synth_pcm(0, 2, 44100 + 100, 44100, 16)),
        ("24b stereo best", Options.best(), 48000, 24, 2, synth_pcm(1, 2, 48000 + 77, 48000, 24)),
        ("16b mono default", Options.default(), 44100, 16, 1, synth_pcm(2, 1, 30000, 44100, 16)),
        ("24b 8ch best", Options.best(), 96000, 24, 8, synth_pcm(4, 8, 12000, 96000, 24)),
        ("16b stereo fast-corr", Options.default().fast_channel_correlation(True), 44100, 16, 2, synth_pcm(7, 2, 30000, 44100, 16)),
        ("8b stereo bs 33", Options.best().block_size(33), 44100, 8, 2, synth_pcm(8, 2, 3000, 44100, 8)),
    ]
    # 1/2/4: generic analyze / lpc / pack; 8: k_pack2 instead of k_pack3; 16: k_analyze instead of k_analyze3;
    # 32: k_lpc2 instead of k_lpc3 / k_lpc4; 56: all second-generation register-tiled kernels; 63: everything generic;
    # 2048: k_lpc3 (four lanes per candidate) instead of k_lpc4 (a lane per candidate); 4096: k_lpc4 per launch group
    for mask in ("7", "1", "2", "4", "8", "16", "32", "56", "63", "2048", "4096"):
        eng.set_option("legacy", int(mask))
        for label, opt, rate, bps, ch, x in cases:
            check(eng, fo, opt, rate, bps, ch, x, f"legacy={mask} {label}")
    eng.set_option("legacy", 0)
    # the single-kernel stereo frame encoder (k_frame4: analysis + decision + look-back placement + packing in one CTA)
    eng.set_option("fused", 1)
    try:
        for label, opt, rate, bps, ch, x in cases:
            check(eng, fo, opt, rate, bps, ch, x, f"fused {label}")
        big = synth_pcm(9, 2, 4096 * 700 + 1234, 48000, 24)     # many frames: the look-back walks over whole waves of CTAs
        check(eng, fo, Options.best(), 48000, 24, 2, big, "fused 24b stereo best, 701 frames")
        for label, opt, rate, bps, ch, x in (("wasted bits", Options.best(), 48000, 24, 2, synth_pcm(3, 2, 20000, 48000, 16) * 256),
                                            ("noise 24b", Options.best(), 48000, 24, 2, np.random.default_rng(3).integers(-(1 << 23), 1 << 23, (9000, 2)).astype(np.int32)),
                                            ("silence", Options.default(), 44100, 16, 2, np.zeros((10000, 2), dtype=np.int32)),
                                            ("bs 4095", Options.best().block_size(4095), 44100, 16, 2, synth_pcm(5, 2, 30000, 44100, 16))):
            check(eng, fo, opt, rate, bps, ch, x, f"fused {label}")
    finally:
        eng.set_option("fused", 0)


def test_frame_kernels_edge_shapes(eng, fo):
    """Shapes that steer the CTA-per-frame kernels (k_lpc3 / k_analyze3 / k_pack3) through their rare paths: PCM that is
    not 16-byte aligned per block (k_lpc3 falls back to k_lpc2), wasted bits (second pass), residuals that overflow the
    int16 copy (FIR recomputed in pass 2), escapes / verbatim / constant subframes, short final blocks, odd block sizes."""
    from flac_codec_b200 import Options, _abi

    rng = np.random.default_rng(5)
    # (a) segment that starts at an odd PCM frame: blocks are not 16-byte aligned
    x = synth_pcm(11, 2, 30000, 48000, 24)
    raw = np.frombuffer(fo.samples_to_bytes(x.reshape(-1), 3), dtype=np.uint8)
    data, sizes, total = eng.encode(Options.best(), 48000, 24, 2, raw, raw.nbytes, _abi.PCM_BYTES_LE, [(1, 30000 - 1, 0)])
    ref, ref_sizes = fo.encode_frames_only(fo.options("best"), 48000, 24, 2, x[1:].reshape(-1))
    assert data.tobytes() == ref and sizes.tolist() == ref_sizes.tolist()
    # (b) wasted bits in one channel only, and in mid/side
    y = synth_pcm(12, 2, 20000, 48000, 24)
    y[:, 0] = (y[:, 0] >> 5) << 5
    check(eng, fo, Options.best(), 48000, 24, 2, y, "wasted left")
    y[:, 1] = (y[:, 1] >> 3) << 3
    check(eng, fo, Options.best(), 48000, 24, 2, y, "wasted both")
    # (c) loud noise and full-scale square bursts: residuals beyond int16, escapes and VERBATIM
    z = rng.integers(-(1 << 23), 1 << 23, size=(3 * 4096 + 100, 2), dtype=np.int64).astype(np.int32)
    z[4096:8192] = np.where((np.arange(4096) // 37) % 2 == 0, (1 << 23) - 1, -(1 << 23))[:, None]
    z[9000:9400] = 0
    check(eng, fo, Options.best(), 48000, 24, 2, z, "noise + square")
    check(eng, fo, Options.default(), 44100, 16, 2, (z >> 8), "noise + square 16-bit")
    # (d) silence, then a frame boundary inside a tone; block sizes that are not multiples of 16 or of 2^p
    w = synth_pcm(13, 2, 11111, 44100, 16)
    w[:5000] = 0
    for bs in (4096, 4095, 1000, 576, 17):
        check(eng, fo, Options.best().block_size(bs), 44100, 16, 2, w, f"block {bs}")
    # (e) 3 and 5 channels (independent mode, odd channel group), big-endian bytes
    for ch in (3, 5):
        v = synth_pcm(14, ch, 9000, 48000, 24)
        check(eng, fo, Options.best(), 48000, 24, ch, v, f"{ch} channels")
        check(eng, fo, Options.best(), 48000, 24, ch, v, f"{ch} channels BE", pcm_kind=_abi.PCM_BYTES_BE)
