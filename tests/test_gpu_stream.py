"""GPU parity of the stream-level facades (include/flacb200_stream.h, flac_codec_b200/stream.py): whole .flac
files written through FlacByteWriter / FlacSampleWriter / FlacChannelWriter must equal the oracle's
FlacSampleWriter restatement (fo_encode_stream: metadata blocks, seek table, MD5, frames) byte for byte, whatever
the write granularity and launch size; readers must return the reference fixtures' PCM (MD5 pinned by
tests/seek.rs:29-31 and STREAMINFO) and reproduce the reference's error behaviour."""
import hashlib
import io

import numpy as np
import pytest

from flacb200_testutil import ref_file, synth_pcm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fo():
    from oracle import oracle

    return oracle


@pytest.fixture(scope="module")
def st():
    from flac_codec_b200 import stream

    return stream


def _options(preset, padding="default", seektable="default", block=None):
    from flac_codec_b200 import Options

    o = Options(preset)
    if block:
        o.block_size(block)
    if padding is None:
        o.no_padding()
    elif padding != "default":
        o.with_padding(padding)
    if seektable is None:
        o.no_seektable()
    elif seektable != "default":
        (o.seektable_seconds if seektable[0] == "seconds" else o.seektable_frames)(seektable[1])
    return o


def _fo_options(fo, preset, padding="default", seektable="default", block=None):
    kw = {}
    if block:
        kw["block_size"] = block
    if padding is None:
        kw["padding"] = -1
    elif padding != "default":
        kw["padding"] = padding if padding > 0 else -1     # Options::padding(0) removes the block (src/encode.rs:1493)
    if seektable is None:
        kw["seektable_kind"] = 0
    elif seektable != "default":
        kw["seektable_kind"] = 1 if seektable[0] == "seconds" else 2
        kw["seektable_n"] = seektable[1]
    return fo.options(preset, **kw)


CASES = [
    # preset, rate, bps, ch, n, total_known, padding, seektable, block, launch_frames, write chunk (pcm frames)
    ("default", 44100, 16, 2, 44100 * 3 + 17, True, "default", "default", None, 0, 10000),
    ("default", 44100, 16, 2, 44100 * 3 + 17, False, "default", "default", None, 7, 4096),
    ("best", 48000, 24, 2, 48000 * 2 + 1, True, "default", ("seconds", 1), None, 5, 777),
    ("best", 48000, 24, 2, 48000 * 2 + 1, False, 64, ("seconds", 1), None, 3, 48000 * 2 + 1),     # padding too small for the table
    ("fast", 8000, 8, 1, 30000, True, None, ("frames", 3), None, 4, 1),                            # byte-at-a-time style writes
    ("fast", 8000, 8, 1, 30000, False, None, ("frames", 3), None, 4, 999),                        # no padding, no placeholder: no table
    ("default", 96000, 24, 8, 96000 // 2 + 5, True, 1000, None, 1024, 2, 5000),
    ("default", 44100, 16, 2, 4096 * 6, True, "default", ("frames", 2), None, 2, 4096),           # exact multiple of the block size
    ("best", 192000, 32, 2, 40000, False, "default", ("seconds", 1), 4608, 0, 12345),
    ("default", 22050, 12, 3, 50000, True, 0, "default", 576, 16, 3333),
]


@pytest.mark.parametrize("preset,rate,bps,ch,n,known,padding,seektable,block,launch,chunk", CASES)
def test_sample_writer_files_equal_oracle(fo, st, preset, rate, bps, ch, n, known, padding, seektable, block, launch, chunk):
    x = synth_pcm(1, ch, n, rate, bps).reshape(-1)
    ref, ref_sizes = fo.encode_stream(_fo_options(fo, preset, padding, seektable, block), rate, bps, ch, x, total_known=known)
    sink = io.BytesIO()
    w = st.FlacSampleWriter(sink, _options(preset, padding, seektable, block), rate, bps, ch, x.size if known else None,
                            launch_frames=launch)
    for i in range(0, n, chunk):
        w.write(x[i * ch:(i + chunk) * ch])
    w.finalize()
    s = w.stats()
    w.close()
    got = sink.getvalue()
    assert len(got) == len(ref)
    assert got == ref
    assert s.frames_written == len(ref_sizes) and s.pcm_frames_written == n
    B = (bps + 7) // 8
    assert bytes(s.md5) == hashlib.md5(fo.samples_to_bytes(x, B)).digest()
    # and the file decodes to the input through the oracle's pinned decoder, MD5 included
    y, si, md5 = fo.decode_stream(got, want_md5=True)
    assert np.array_equal(y, x) and bytes(si.md5) == md5


@pytest.mark.parametrize("endian", ["little", "big"])
@pytest.mark.parametrize("bps", [8, 16, 24, 32])
def test_byte_writer_and_channel_writer(fo, st, endian, bps):
    rate, ch, n = 32000, 2, 20000
    x = synth_pcm(2, ch, n, rate, bps)
    ref, _ = fo.encode_stream(fo.options("default"), rate, bps, ch, x.reshape(-1), total_known=True)
    B = (bps + 7) // 8
    raw = fo.samples_to_bytes(x.reshape(-1), B, big_endian=(endian == "big"))
    sink = io.BytesIO()
    from flac_codec_b200 import Options

    with st.FlacByteWriter(sink, Options.default(), rate, bps, ch, len(raw), endian=endian, launch_frames=2) as w:
        for i in range(0, len(raw), 7001):      # splits samples across writes
            assert w.write(raw[i:i + 7001]) == len(raw[i:i + 7001])
    assert sink.getvalue() == ref
    sink = io.BytesIO()
    with st.FlacChannelWriter(sink, Options.default(), rate, bps, ch, n, launch_frames=3) as w:
        for i in range(0, n, 2500):
            w.write([x[i:i + 2500, c] for c in range(ch)])
    assert sink.getvalue() == ref


def test_writer_errors(fo, st):
    from flac_codec_b200 import Options
    from flac_codec_b200._abi import FlacB200Error

    x = synth_pcm(0, 2, 10000, 44100, 16).reshape(-1)
    # more samples than announced: Encoder::encode -> ExcessiveTotalSamples (src/encode.rs:2006-2011)
    w = st.FlacSampleWriter(io.BytesIO(), Options.default(), 44100, 16, 2, 4096 * 2, launch_frames=64)
    with pytest.raises(FlacB200Error) as e:
        w.write(x)
    assert e.value.code == 57
    w.close()
    # fewer: finalize -> SampleCountMismatch (:2080-2084)
    w = st.FlacSampleWriter(io.BytesIO(), Options.default(), 44100, 16, 2, x.size + 2)
    w.write(x)
    with pytest.raises(FlacB200Error) as e:
        w.finalize()
    assert e.value.code == 59
    w.close()
    # nothing written: NoSamples (:2089)
    w = st.FlacSampleWriter(io.BytesIO(), Options.default(), 44100, 16, 2, None)
    with pytest.raises(FlacB200Error) as e:
        w.finalize()
    assert e.value.code == 58
    w.close()
    with pytest.raises(FlacB200Error) as e:
        st.FlacSampleWriter(io.BytesIO(), Options.default(), 44100, 16, 2, 7)
    assert e.value.code == 61
    w = st.FlacChannelWriter(io.BytesIO(), Options.default(), 44100, 16, 2, None)
    with pytest.raises(FlacB200Error) as e:
        w.write([x[:10]])
    assert e.value.code == 64
    with pytest.raises(FlacB200Error) as e:
        w.write([x[:10], x[:9]])
    assert e.value.code == 65
    w.close()


def test_stream_writer_subset_frames(fo, st):
    from flac_codec_b200 import Options
    from flac_codec_b200._abi import FlacB200Error

    sink = io.BytesIO()
    w = st.FlacStreamWriter(sink, Options.default())
    x = synth_pcm(3, 2, 3000, 44100, 16)
    w.write(44100, 2, 16, x[:1000].reshape(-1))
    w.write_cdda(x[1000:].reshape(-1))
    ref = b""
    for k, blk in enumerate((x[:1000], x[1000:])):
        ref += fo.encode_frame(fo.options("default", block_size=len(blk)), 44100, 16, blk.T, frame_number=k, subset=True)
    assert sink.getvalue() == ref
    r = st.FlacStreamReader(sink.getvalue())
    a, rate, ch, bps = r.read()
    assert (rate, ch, bps) == (44100, 2, 16) and np.array_equal(a.reshape(-1, 2), x[:1000])
    b, rate, ch, bps = r.read()
    assert (rate, ch, bps) == (44100, 2, 16) and np.array_equal(b.reshape(-1, 2), x[1000:])
    with pytest.raises(FlacB200Error) as e:
        r.read()
    assert e.value.code == 1     # Io: "eof looking for frame sync" (src/decode.rs:1198)
    r.close()
    with pytest.raises(FlacB200Error) as e:
        w.write(44100, 2, 17, x[:10].reshape(-1))     # NonSubsetBitsPerSample (:1134)
    assert e.value.code == 28
    with pytest.raises(FlacB200Error) as e:
        w.write(44100, 2, 16, x.reshape(-1)[:11])     # SamplesNotDivisibleByChannels (:1108)
    assert e.value.code == 61


@pytest.mark.parametrize("name,md5", [("sine.flac", "831671b807f97051301e01d68b5c54b3"), ("all-frames.flac", None),
                                      ("cuesheet.flac", None), ("seektable.flac", None)])
def test_readers_on_reference_fixtures(fo, st, name, md5):
    flac = ref_file(name)
    y, si, oracle_md5 = fo.decode_stream(flac, want_md5=True)
    r = st.FlacSampleReader(io.BytesIO(flac))
    assert (r.channel_count(), r.sample_rate(), r.bits_per_sample(), r.total_samples()) == (si.channels, si.sample_rate, si.bps,
                                                                                             si.total_samples or None)
    got = r.read_to_end()
    assert np.array_equal(got, y)
    status, sum_ = r.verify()
    assert sum_ == oracle_md5
    if md5:
        assert sum_.hex() == md5
    assert status == ("MD5Match" if any(bytes(si.md5)) else "NoMD5")
    # seek (tests/seek.rs): position in inter-channel samples, then the same samples as the full decode
    total = len(y) // si.channels
    for pos in (0, 1, total // 3, total - 1, total):
        r.seek(pos)
        a = r.read(4096)
        assert np.array_equal(a, y[pos * si.channels:pos * si.channels + 4096])
    from flac_codec_b200._abi import FlacB200Error

    with pytest.raises(FlacB200Error) as e:
        r.seek(total + 1)
    assert e.value.code == 37     # InvalidSeek
    r.close()
    B = (si.bps + 7) // 8
    for endian in ("little", "big"):
        rb = st.FlacByteReader(flac, endian=endian)
        assert rb.read() == fo.samples_to_bytes(y, B, big_endian=(endian == "big"))
        assert rb.decoded_len() == (len(y) * B if si.total_samples else None)
        rb.close()
    assert st.verify(flac) == ("MD5Match" if any(bytes(si.md5)) else "NoMD5")


def test_reader_surfaces_corruption_after_the_good_frames(fo, st):
    from flac_codec_b200 import Options
    from flac_codec_b200._abi import FlacB200Error

    x = synth_pcm(5, 2, 4096 * 6, 44100, 16).reshape(-1)
    flac, sizes = fo.encode_stream(fo.options("default"), 44100, 16, 2, x, total_known=True)
    si = fo.read_streaminfo(flac)
    bad = bytearray(flac)
    bad[si.frames_start + int(sizes[:3].sum()) + int(sizes[3]) // 2] ^= 0x10     # inside frame 3
    r = st.FlacSampleReader(bytes(bad))
    got = []
    with pytest.raises(FlacB200Error) as e:
        while True:
            a = r.read(5000)
            if a.size == 0:
                break
            got.append(a.copy())
    assert e.value.code in (39, 40)     # Crc8Mismatch / Crc16Mismatch, as the oracle reports
    assert np.array_equal(np.concatenate(got), x[:3 * 4096 * 2])
    with pytest.raises(FlacB200Error):
        r.verify()
    r.close()
    # a different MD5 in STREAMINFO: Verified::MD5Mismatch
    wrong = bytearray(flac)
    wrong[4 + 4 + 18] ^= 0xFF
    assert st.verify(bytes(wrong)) == "MD5Mismatch"


def test_cpp_host_front_end_wav2flac_roundtrip(fo, tmp_path):
    """The C++ facades (flac_codec_b200/host/flacb200.hpp) through the wav2flac/flac2wav front end
    (examples/wav2flac.rs, examples/flac2wav.rs): file equals the oracle's FlacByteWriter restatement; WAV survives."""
    import struct
    import subprocess

    from flac_codec_b200 import build

    exe = build.build_host_tools()
    rate, bps, ch, n = 44100, 16, 2, 44100 * 2 + 123
    x = synth_pcm(7, ch, n, rate, bps).reshape(-1)
    raw = fo.samples_to_bytes(x, 2)
    wav = (b"RIFF" + struct.pack("<I", 36 + len(raw)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, ch, rate, rate * ch * 2, ch * 2, bps)
           + b"data" + struct.pack("<I", len(raw)) + raw)
    (tmp_path / "in.wav").write_bytes(wav)
    subprocess.check_call([exe, "encode", str(tmp_path / "in.wav"), str(tmp_path / "out.flac")])
    ref, _ = fo.encode_stream(fo.options("default"), rate, bps, ch, x, total_known=True)
    assert (tmp_path / "out.flac").read_bytes() == ref
    subprocess.check_call([exe, "decode", str(tmp_path / "out.flac"), str(tmp_path / "back.wav")])
    assert (tmp_path / "back.wav").read_bytes() == wav


# ---- the reader handle: windows, feeding, seek through the SEEKTABLE, fill_buf / consume, channels ----
def _long_stream(fo, seconds=20, rate=44100, bps=16, ch=2, **optkw):
    x = synth_pcm(21, ch, rate * seconds + 777, rate, bps).reshape(-1)
    flac, sizes = fo.encode_stream(fo.options("default", **optkw), rate, bps, ch, x, total_known=True)
    return x, flac, sizes


def test_reader_windows_are_invisible(fo, st):
    """A stream decoded in many small windows (frames cut by the window ends, tiny PCM budgets) is the stream."""
    x, flac, sizes = _long_stream(fo)
    for wb, wp in ((1 << 16, 0), (5000, 0), (1 << 20, 4096), (12345, 9000)):
        r = st.FlacSampleReader(flac, window_bytes=wb, window_pcm_frames=wp)
        assert np.array_equal(r.read_to_end(), x), (wb, wp)
        assert r.verify()[0] == "MD5Match"
        r.close()
    rb = st.FlacByteReader(flac, window_bytes=70000)
    assert rb.read() == fo.samples_to_bytes(x, 2)
    rb.close()
    # the same through caller-owned buffers (no intermediate copies); a buffer larger than the stream is filled to its end
    rb = st.FlacByteReader(flac, window_bytes=70000)
    out = np.zeros(x.size * 2 + 100, dtype=np.uint8)
    assert rb.readinto(out) == x.size * 2 and out[: x.size * 2].tobytes() == fo.samples_to_bytes(x, 2)
    rb.close()
    rs = st.FlacSampleReader(flac, window_bytes=50000)
    y = np.zeros(x.size // 2, dtype=np.int32)
    assert rs.readinto(y) == y.size and np.array_equal(y, x[: y.size])
    assert rs.readinto(y) == x.size - y.size and np.array_equal(y[: x.size - y.size], x[y.size:])
    rs.close()


def test_reader_fed_in_chunks_like_a_plain_read(fo, st):
    """FlacSampleReader::new(R: Read): the file arrives in chunks, decoded bytes are dropped, seeking is refused."""
    from flac_codec_b200._abi import FlacB200Error

    x, flac, _ = _long_stream(fo, seconds=8)
    for chunk in (1000, 65536, 1 << 22):
        r = st.FlacSampleReader(io.BytesIO(flac), streaming=True, chunk=chunk, window_bytes=1 << 17)
        assert r.total_samples() == x.size // 2 and r.sample_rate() == 44100
        got = []
        while True:
            a = r.read(50000)
            if a.size == 0:
                break
            got.append(a.copy())
        assert np.array_equal(np.concatenate(got), x), chunk
        with pytest.raises(FlacB200Error):
            r.seek(0)
        r.close()
    assert st.FlacByteReader(io.BytesIO(flac), streaming=True, chunk=4096).verify()[0] == "MD5Match"
    # a source that ends inside a frame: the frames in front are delivered, then Io
    r = st.FlacSampleReader(io.BytesIO(flac[: len(flac) * 2 // 3]), streaming=True, chunk=30000)
    got = []
    with pytest.raises(FlacB200Error) as e:
        while True:
            a = r.read(1 << 20)
            if a.size == 0:
                break
            got.append(a.copy())
    assert e.value.code == 1
    n = sum(a.size for a in got)
    assert n > 0 and n % (4096 * 2) == 0 and np.array_equal(np.concatenate(got), x[:n])
    r.close()


def test_seek_uses_the_seektable(fo, st):
    """Decoder::seek (src/decode.rs:1452-1491): jump to the last seek point <= sample, decode and skip from there.  With a
    window of one frame the number of frames decoded after a seek shows that the jump happened."""
    from flac_codec_b200._abi import FlacB200Error

    x, flac, sizes = _long_stream(fo, seconds=35)     # default seek table: every 10 s
    total = x.size // 2
    r = st.FlacSampleReader(flac)
    pts = [p for p in r.seektable() if not p[3]]
    assert len(pts) == 4 and pts[1][0] > 0
    for pos in (0, 5, 4096 * 3 + 17, pts[1][0] - 1, pts[1][0], pts[2][0] + 4096 * 5 + 1, total - 1, total):
        r.seek(pos)
        assert np.array_equal(r.read(6000), x[pos * 2:pos * 2 + 6000]), pos
    with pytest.raises(FlacB200Error) as e:
        r.seek(total + 1)
    assert e.value.code == 37
    r.close()
    # the same file with its seek points blanked out: still correct, from the start of the stream
    si = fo.read_streaminfo(flac)
    blank = bytearray(flac)
    at = 4 + 4 + 34 + 4
    for i in range(si.n_seekpoints if hasattr(si, "n_seekpoints") else len(pts)):
        blank[at + 18 * i: at + 18 * i + 8] = b"\xff" * 8
    r = st.FlacSampleReader(bytes(blank))
    assert all(p[3] for p in r.seektable())
    r.seek(pts[2][0] + 100)
    assert np.array_equal(r.read(1000), x[(pts[2][0] + 100) * 2:(pts[2][0] + 100) * 2 + 1000])
    r.close()
    # FlacByteReader's io::Seek, in bytes (src/decode.rs:715-820)
    rb = st.FlacByteReader(flac)
    raw = fo.samples_to_bytes(x, 2)
    for off in (0, 4 * 1000, 4 * (pts[1][0] + 3), len(raw) - 8, 4 * 777 + 1, 4 * (pts[2][0] + 9) + 3):   # the last two: inside a PCM frame
        assert rb.seek_bytes(off) == off
        assert rb.read(4000) == raw[off:off + 4000]
        assert rb.read(3) == raw[off + 4000:off + 4003] and rb.read(6) == raw[off + 4003:off + 4009]   # reads are byte-granular
    # SeekFrom::End: the reference takes total_samples (PCM frames, not bytes) as the end position (src/decode.rs:766-768)
    total = len(raw) // 4
    assert rb.seek_bytes(-400, 2) == total - 400 and rb.read(100) == raw[total - 400:total - 300]
    rb.close()
    # FlacSampleReader::read (:417) hands out any number of samples, also half a stereo pair
    rs = st.FlacSampleReader(flac)
    assert np.array_equal(rs.read(1), x[:1]) and np.array_equal(rs.read(1), x[1:2]) and np.array_equal(rs.read(5), x[2:7])
    buf = rs.fill_buf()
    assert np.array_equal(buf[:9], x[7:16])
    rs.consume(3)
    assert np.array_equal(rs.read(4), x[10:14])
    rs.seek(100)
    assert np.array_equal(rs.read(3), x[200:203])
    rs.close()


def test_fill_buf_consume_and_channel_reader(fo, st):
    """fill_buf hands out one frame at a time (FlacSampleReader :466, FlacChannelReader :917); block sizes vary in
    all-frames.flac."""
    for name in ("all-frames.flac", "sine.flac"):
        flac = ref_file(name)
        y, si = fo.decode_stream(flac)
        ch = si.channels
        # frame sizes as the serial reader sees them
        sizes, p, cur = [], si.frames_start, 0
        while cur < si.total_samples:
            _, h, used = fo.decode_frame(flac[p:], si, si.total_samples - cur)
            sizes.append(h.block_size)
            p += used
            cur += h.block_size
        r = st.FlacSampleReader(flac)
        pos = 0
        for k, bs in enumerate(sizes):
            buf = r.fill_buf()
            assert buf.size == bs * ch, (name, k)
            assert np.array_equal(buf, y[pos:pos + bs * ch])
            if bs > 1:     # partial consume: the rest of the same frame comes back
                r.consume(ch)
                assert np.array_equal(r.fill_buf(), y[pos + ch:pos + bs * ch])
                r.consume((bs - 1) * ch)
            else:
                r.consume(ch)
            pos += bs * ch
        assert r.fill_buf().size == 0
        r.close()
        rc = st.FlacChannelReader(flac)
        pos = 0
        for bs in sizes:
            chans = rc.fill_buf()
            assert len(chans) == ch and all(c.size == bs for c in chans)
            for c in range(ch):
                assert np.array_equal(chans[c], y[pos + c:pos + bs * ch:ch])
            rc.consume(bs)
            pos += bs * ch
        assert all(c.size == 0 for c in rc.fill_buf())
        rc.seek(sizes[0] + 1)
        chans = rc.fill_buf()
        assert np.array_equal(chans[0], y[(sizes[0] + 1) * ch:(sizes[0] + sizes[1]) * ch:ch])
        rc.close()
    # FlacSampleIterator (:667)
    flac = ref_file("all-frames.flac")
    y, si = fo.decode_stream(flac)
    assert list(st.FlacSampleReader(flac)) == y.tolist()


def test_stream_reader_takes_its_parameters_from_the_frames(fo, st):
    """FlacStreamReader::read (src/decode.rs:1175-1240): no metadata, no parameters from the caller; rate, channels and
    bits per sample change from frame to frame; junk between frames is skipped by the sync scan."""
    from flac_codec_b200 import Options
    from flac_codec_b200._abi import FlacB200Error

    sink = io.BytesIO()
    w = st.FlacStreamWriter(sink, Options.default())
    plan = [(44100, 2, 16, 1000), (44100, 2, 16, 4096), (48000, 2, 16, 777), (96000, 1, 24, 2000), (96000, 1, 24, 2000), (8000, 4, 8, 333)]
    blocks = []
    for k, (rate, ch, bps, n) in enumerate(plan):
        x = synth_pcm(40 + k, ch, n, rate, bps)
        blocks.append(x)
        w.write(rate, ch, bps, x.reshape(-1))
    data = sink.getvalue()
    for chunk in (1 << 20, 997):
        r = st.FlacStreamReader(io.BytesIO(data), chunk=chunk)
        for (rate, ch, bps, n), x in zip(plan, blocks):
            a, grate, gch, gbps = r.read()
            assert (grate, gch, gbps) == (rate, ch, bps)
            assert np.array_equal(a.reshape(-1, ch), x)
        with pytest.raises(FlacB200Error) as e:
            r.read()
        assert e.value.code == 1
        r.close()
    # junk in front and between two frames
    f0 = fo.encode_frame(fo.options("default", block_size=500), 44100, 16, blocks[0][:500].T, frame_number=0, subset=True)
    f1 = fo.encode_frame(fo.options("default", block_size=300), 44100, 16, blocks[0][500:800].T, frame_number=1, subset=True)
    r = st.FlacStreamReader(b"\x00\x01\xff\x00junk" + f0 + b"\xff\xf8\x00garbage\xff" + f1)
    a, *_ = r.read()
    b, *_ = r.read()
    assert np.array_equal(a.reshape(-1, 2), blocks[0][:500]) and np.array_equal(b.reshape(-1, 2), blocks[0][500:800])
    r.close()


def test_fed_reader_over_a_seekable_source_holds_nothing(fo, st):
    """new_seekable without a file image: the handle asks the caller to reposition its source (FLACB200_NEED_SEEK) and is fed
    from the seek point on."""
    x, flac, _ = _long_stream(fo, seconds=35)
    total = x.size // 2
    src = io.BytesIO(flac)
    r = st.FlacSampleReader(src, streaming=True, seekable=True, chunk=50000, window_bytes=1 << 16)
    pts = [p for p in r.seektable() if not p[3]]
    for pos in (pts[2][0] + 5000, 17, pts[1][0], total - 3, 0):
        r.seek(pos)
        assert np.array_equal(r.read(5000), x[pos * 2:pos * 2 + 5000]), pos
    r.seek(0)
    assert r.verify()[0] == "MD5Match"
    r.close()
