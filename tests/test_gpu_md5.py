"""GPU many-stream MD5 (flacb200_md5_batch) against hashlib: the STREAMINFO signature is MD5 over the samples as
little-endian interleaved bytes, ceil(bps / 8) bytes each (update_md5, src/encode.rs:1292-1318)."""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from flac_codec_b200 import Engine

    e = Engine(0)
    yield e
    e.close()


def le_bytes(x: np.ndarray, nbytes: int) -> bytes:
    """int32 samples -> nbytes little-endian bytes each"""
    b = x.astype("<i4").view(np.uint8).reshape(-1, 4)[:, :nbytes]
    return b.tobytes()


@pytest.mark.parametrize("bps,ch", [(8, 1), (16, 2), (24, 2), (24, 3), (32, 2), (12, 1), (20, 5)])
def test_md5_lengths_and_offsets_packed_le(eng, bps, ch):
    """Message lengths around the block and padding boundaries (0, 1, 55, 56, 63, 64, 65, 119, 120, 128, ...), streams that
    start at every byte alignment inside one buffer."""
    from flac_codec_b200 import _abi

    nb = (bps + 7) // 8
    rng = np.random.default_rng(bps * 10 + ch)
    lens = [0, 1, 2, 3, 4, 5, 7, 9, 10, 11, 16, 21, 22, 31, 32, 33, 43, 64, 100, 127, 128, 129, 500, 4097]   # PCM frames
    segs, off = [], 0
    for n in lens:
        off += int(rng.integers(0, 4))   # gaps move the start alignment around
        segs.append((off, n))
        off += n
    total = off + 3
    x = rng.integers(-(1 << (bps - 1)), 1 << (bps - 1), size=(total, ch), dtype=np.int64).astype(np.int32)
    raw = np.frombuffer(le_bytes(x.reshape(-1), nb), dtype=np.uint8).copy()
    got = eng.md5(bps, ch, raw, raw.nbytes, _abi.PCM_BYTES_LE, segs)
    for (o, n), d in zip(segs, got):
        want = hashlib.md5(raw[o * ch * nb:(o + n) * ch * nb].tobytes()).digest()
        assert d == want, (bps, ch, o, n)


@pytest.mark.parametrize("bps", [16, 24])
def test_md5_other_layouts(eng, bps):
    """Big-endian bytes and int32 samples hash to the same digest as the little-endian message."""
    from flac_codec_b200 import _abi

    ch, n, nb = 2, 1000, (bps + 7) // 8
    rng = np.random.default_rng(7)
    x = rng.integers(-(1 << (bps - 1)), 1 << (bps - 1), size=(n, ch), dtype=np.int64).astype(np.int32)
    msg = le_bytes(x.reshape(-1), nb)
    segs = [(0, n), (10, 333), (999, 1), (500, 0)]
    want = [hashlib.md5(msg[o * ch * nb:(o + k) * ch * nb]).digest() for o, k in segs]
    be = np.frombuffer(msg, dtype=np.uint8).reshape(-1, nb)[:, ::-1].copy().reshape(-1)
    assert eng.md5(bps, ch, be, be.nbytes, _abi.PCM_BYTES_BE, segs) == want
    inter = x.reshape(-1).copy()
    assert eng.md5(bps, ch, inter, inter.nbytes, _abi.PCM_I32_INTERLEAVED, segs) == want
    planar = np.ascontiguousarray(x.T).reshape(-1)
    assert eng.md5(bps, ch, planar, planar.nbytes, _abi.PCM_I32_PLANAR, segs, planar_stride=n) == want


def test_md5_matches_streaminfo_of_written_file(eng):
    """The digest of a track equals the MD5 the writer facade stores in STREAMINFO (and the oracle's)."""
    import io

    from flac_codec_b200 import Options, _abi, stream
    from flacb200_testutil import synth_pcm
    from oracle import oracle as fo

    rate, bps, ch = 48000, 24, 2
    x = synth_pcm(3, ch, 30000, rate, bps)
    raw = np.frombuffer(fo.samples_to_bytes(x.reshape(-1), 3), dtype=np.uint8).copy()
    sink = io.BytesIO()
    w = stream.FlacByteWriter(sink, Options.best(), rate, bps, ch, raw.nbytes, engine=eng)
    w.write(raw.tobytes())
    w.finalize()
    w.close()
    si = fo.read_streaminfo(sink.getvalue())
    got = eng.md5(bps, ch, raw, raw.nbytes, _abi.PCM_BYTES_LE, [(0, x.shape[0])])[0]
    assert got == bytes(si.md5) == hashlib.md5(raw.tobytes()).digest()


def test_md5_device_resident_batch(eng):
    """Many tracks resident on the device (the bench shape, shortened): one digest per track."""
    from flac_codec_b200 import _abi

    rate, bps, ch, ntr, n = 48000, 24, 2, 64, 48000
    nbytes = ntr * n * ch * 3
    d = eng.device_alloc(nbytes)
    eng.synth_pcm(d, 0, ntr, n, ch, rate, bps)
    host = np.zeros(nbytes, dtype=np.uint8)
    eng.memcpy(host, d, nbytes, 2)
    got = eng.md5(bps, ch, d, nbytes, _abi.PCM_BYTES_LE, [(t * n, n) for t in range(ntr)], pcm_location=_abi.DEVICE)
    eng.device_free(d)
    per = n * ch * 3
    for t in range(ntr):
        assert got[t] == hashlib.md5(host[t * per:(t + 1) * per].tobytes()).digest()
