"""Whole-file batches through the C ABI (flacb200_encode_batch / flacb200_decode_batch): every file equals the oracle's
FlacByteWriter restatement (metadata, seek table, MD5, frames), also when the tracks are dealt to several device workers."""
import hashlib

import numpy as np
import pytest

from flacb200_testutil import ref_file, synth_pcm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fo():
    from oracle import oracle

    return oracle


def _tracks(fo):
    from flac_codec_b200 import _abi

    specs = [(44100, 16, 2, 44100 * 3 + 11), (44100, 16, 2, 50000), (48000, 24, 2, 48000 * 2), (44100, 16, 2, 4096 * 4), (96000, 24, 8, 9000),
             (48000, 24, 2, 12345), (44100, 16, 1, 30000), (44100, 16, 2, 15), (192000, 32, 2, 20000)]
    out = []
    for k, (rate, bps, ch, n) in enumerate(specs):
        x = synth_pcm(60 + k, ch, n, rate, bps).reshape(-1)
        B = (bps + 7) // 8
        kind = (_abi.PCM_BYTES_LE, _abi.PCM_BYTES_BE, _abi.PCM_I32_INTERLEAVED)[k % 3]
        if kind == _abi.PCM_I32_INTERLEAVED:
            pcm = x.astype(np.int32)
        else:
            pcm = np.frombuffer(fo.samples_to_bytes(x, B, big_endian=(kind == _abi.PCM_BYTES_BE)), dtype=np.uint8).copy()
        out.append((x, (pcm, n, rate, bps, ch, kind)))
    return out


@pytest.mark.parametrize("preset", ["default", "best"])
@pytest.mark.parametrize("devices", [[0], [0, 0, 0]])
def test_encode_batch_files_equal_the_oracle(fo, preset, devices):
    from flac_codec_b200 import Options
    from flac_codec_b200.batch import encode_files

    tr = _tracks(fo)
    got = encode_files([t for _, t in tr], Options(preset), devices=devices)
    for (x, (pcm, n, rate, bps, ch, kind)), (data, status, md5) in zip(tr, got):
        assert status == 0, (rate, bps, ch, n, status)
        ref, _ = fo.encode_stream(fo.options(preset), rate, bps, ch, x, total_known=True)
        assert data == ref, (rate, bps, ch, n)
        assert md5 == hashlib.md5(fo.samples_to_bytes(x, (bps + 7) // 8)).digest()


def test_encode_batch_errors_are_per_track(fo):
    from flac_codec_b200 import Options, _abi
    from flac_codec_b200.batch import encode_files

    x = synth_pcm(1, 2, 5000, 44100, 16).reshape(-1)
    raw = np.frombuffer(fo.samples_to_bytes(x, 2), dtype=np.uint8).copy()
    got = encode_files([(raw, 5000, 44100, 16, 2, _abi.PCM_BYTES_LE), (raw, 0, 44100, 16, 2, _abi.PCM_BYTES_LE),
                        (raw, 2500, 44100, 16, 9, _abi.PCM_BYTES_LE), (raw, 5000, 1 << 20, 16, 2, _abi.PCM_BYTES_LE)], Options.default())
    assert [s for _, s, _ in got] == [0, 58, 30, 26]     # ok, NoSamples, ExcessiveChannels, InvalidSampleRate
    assert got[0][0] == fo.encode_stream(fo.options("default"), 44100, 16, 2, x, total_known=True)[0]


@pytest.mark.parametrize("devices", [[0], [0, 0]])
def test_decode_batch_roundtrip_verify_and_damage(fo, devices):
    from flac_codec_b200 import _abi
    from flac_codec_b200.batch import decode_files

    tr = _tracks(fo)
    flacs, want = [], []
    for x, (pcm, n, rate, bps, ch, kind) in tr:
        flacs.append(fo.encode_stream(fo.options("default"), rate, bps, ch, x, total_known=True)[0])
        want.append((x, bps))
    for name in ("sine.flac", "all-frames.flac"):
        f = ref_file(name)
        flacs.append(f)
        y, si = fo.decode_stream(f)
        want.append((y, si.bps))
    bad = bytearray(flacs[2])
    bad[len(bad) // 2] ^= 0x04
    wrong_md5 = bytearray(flacs[3])
    wrong_md5[4 + 4 + 18] ^= 0xFF
    flacs += [bytes(bad), bytes(wrong_md5), b"not a flac file at all"]
    for kind in (_abi.PCM_BYTES_LE, _abi.PCM_I32_INTERLEAVED):
        got = decode_files(flacs, kind, verify=True, devices=devices)
        for (x, bps), (pcm, status, verified, info) in zip(want, got):
            assert status == 0 and verified in (0, 2)
            if kind == _abi.PCM_I32_INTERLEAVED:
                assert np.array_equal(np.frombuffer(pcm, dtype=np.int32), x)
            else:
                assert pcm == fo.samples_to_bytes(x, (bps + 7) // 8)
        assert got[len(want)][1] in (39, 40) and got[len(want)][0] is None      # the damaged stream, and only it
        assert got[len(want) + 1][1] == 0 and got[len(want) + 1][2] == 1       # MD5Mismatch
        assert got[len(want) + 2][1] == 3                                       # MissingFlacTag
