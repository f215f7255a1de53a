"""CPU-side checks of the stream-level host layer (include/flacb200_stream.h): MD5, metadata walk of the
reference's own .flac fixtures against the oracle's, the writer's initial metadata blocks, the facade argument
checks (src/encode.rs:170-178, :516-521) -- and that, without a GPU, a writer fails loudly instead of encoding
on the CPU."""
import ctypes as C
import hashlib
import io

import numpy as np
import pytest

from flacb200_testutil import ref_file


@pytest.fixture(scope="module")
def L():
    from flac_codec_b200 import _abi, build

    build.build()
    return _abi.lib()


@pytest.fixture(scope="module")
def fo():
    from oracle import oracle

    oracle.build()
    return oracle


@pytest.mark.parametrize("n", [0, 1, 55, 56, 57, 63, 64, 65, 119, 120, 121, 1000, 65536 + 17])
def test_md5_matches_hashlib(L, n):
    data = np.random.default_rng(n).integers(0, 256, n, dtype=np.uint8)
    out = (C.c_uint8 * 16)()
    L.flacb200_md5(C.c_void_p(data.ctypes.data if n else 0), n, C.byref(out))
    assert bytes(out) == hashlib.md5(data.tobytes()).digest()


@pytest.mark.parametrize("name", ["sine.flac", "all-frames.flac", "cuesheet.flac", "seektable.flac"])
def test_metadata_walk_matches_oracle(L, fo, name):
    from flac_codec_b200 import _abi

    flac = ref_file(name)
    a = np.frombuffer(flac, dtype=np.uint8)
    si = _abi.Streaminfo()
    assert L.flacb200_read_streaminfo(C.c_void_p(a.ctypes.data), a.size, C.byref(si)) == 0
    ref = fo.read_streaminfo(flac)
    assert (si.min_block_size, si.max_block_size, si.min_frame_size, si.max_frame_size, si.sample_rate, si.channels,
            si.bits_per_sample, si.total_samples, bytes(si.md5), si.frames_start) == (
        ref.min_block_size, ref.max_block_size, ref.min_frame_size, ref.max_frame_size, ref.sample_rate, ref.channels,
        ref.bps, ref.total_samples, bytes(ref.md5), ref.frames_start)


def test_seektable_fixture_points(L):
    from flac_codec_b200 import _abi

    flac = ref_file("seektable.flac")
    a = np.frombuffer(flac, dtype=np.uint8)
    h = C.c_void_p()
    assert L.flacb200_reader_open(None, C.c_void_p(a.ctypes.data), a.size, C.byref(h)) == 0
    n = C.c_size_t(0)
    assert L.flacb200_reader_seektable(h, None, 0, C.byref(n)) == 0
    assert n.value > 0
    pts = (_abi.SeekPoint * n.value)()
    assert L.flacb200_reader_seektable(h, pts, n.value, C.byref(n)) == 0
    defined = [p.sample_offset for p in pts if not p.placeholder]
    assert defined == sorted(set(defined)) and defined[0] == 0
    # without a GPU, decoding fails loudly (no CPU fallback)
    import torch

    if not torch.cuda.is_available():
        out = np.zeros(16, dtype=np.int32)
        got = C.c_size_t(0)
        assert L.flacb200_reader_read(h, C.c_void_p(out.ctypes.data), 16, _abi.PCM_I32_INTERLEAVED, C.byref(got)) == -1
    L.flacb200_reader_close(h)


def test_metadata_errors(L):
    from flac_codec_b200 import _abi

    def rc(b):
        a = np.frombuffer(b, dtype=np.uint8)
        si = _abi.Streaminfo()
        return L.flacb200_read_streaminfo(C.c_void_p(a.ctypes.data), a.size, C.byref(si))

    good = ref_file("sine.flac")
    assert rc(good) == 0
    assert rc(b"fLaX" + good[4:]) == 3                       # MissingFlacTag
    assert rc(good[:4] + bytes([0x04]) + good[5:]) == 4      # first block is not STREAMINFO -> MissingStreaminfo
    si_block = good[4:4 + 38]
    not_last = bytes([si_block[0] & 0x7F]) + si_block[1:]
    assert rc(good[:4] + not_last + bytes([0x80]) + si_block[1:]) == 5                   # MultipleStreaminfo
    assert rc(good[:4] + not_last + bytes([0x80 | 50, 0, 0, 0])) == 14                   # ReservedMetadataBlock
    assert rc(good[:4] + not_last + bytes([0x80 | 127, 0, 0, 0])) == 15                  # InvalidMetadataBlock
    assert rc(good[:4] + not_last + bytes([0x80 | 3, 0, 0, 17]) + bytes(17)) == 8        # InvalidSeekTableSize
    two = bytes([3, 0, 0, 18]) + bytes(18) + bytes([0x80 | 3, 0, 0, 18]) + bytes(18)
    assert rc(good[:4] + not_last + two) == 6                                            # MultipleSeekTable


def test_total_checks(L):
    out = C.c_uint64(0)
    assert L.flacb200_total_from_bytes(4096 * 4, 16, 2, C.byref(out)) == 0 and out.value == 4096
    assert L.flacb200_total_from_bytes(4097, 16, 2, C.byref(out)) == 61    # SamplesNotDivisibleByChannels
    assert L.flacb200_total_from_bytes(0, 16, 2, C.byref(out)) == 62       # InvalidTotalBytes
    assert L.flacb200_total_from_bytes(9, 24, 1, C.byref(out)) == 0 and out.value == 3
    assert L.flacb200_total_from_samples(7, 2, C.byref(out)) == 61
    assert L.flacb200_total_from_samples(0, 2, C.byref(out)) == 63         # InvalidTotalSamples
    assert L.flacb200_total_from_samples(8, 2, C.byref(out)) == 0 and out.value == 4


def _open(L, wo, rate, bps, ch, total):
    h = C.c_void_p()
    rc = L.flacb200_writer_open(None, C.byref(wo), rate, bps, ch, total, C.byref(h))
    return rc, h


def test_writer_initial_metadata_and_argument_errors(L):
    from flac_codec_b200 import _abi

    wo = _abi.WriterOptions()
    L.flacb200_writer_options_default(C.byref(wo))
    assert (wo.padding, wo.seektable_kind, wo.seektable_n, wo.frame.block_size) == (4096, 1, 10, 4096)
    assert _open(L, wo, 1 << 20, 16, 2, 0)[0] == 26   # InvalidSampleRate
    assert _open(L, wo, 44100, 16, 9, 0)[0] == 30     # ExcessiveChannels
    assert _open(L, wo, 44100, 33, 2, 0)[0] == 33     # InvalidBitsPerSample
    assert _open(L, wo, 44100, 16, 2, 1 << 36)[0] == 57   # ExcessiveTotalSamples
    # known total: STREAMINFO + placeholder SEEKTABLE (one point per 10 s) + PADDING
    total = 44100 * 25
    rc, h = _open(L, wo, 44100, 16, 2, total)
    assert rc == 0
    p, n = C.POINTER(C.c_uint8)(), C.c_size_t(0)
    assert L.flacb200_writer_header(h, C.byref(p), C.byref(n)) == 0
    hdr = C.string_at(p, n.value)
    assert hdr[:4] == b"fLaC" and hdr[4] == 0 and hdr[5:8] == bytes([0, 0, 34])
    si = _abi.Streaminfo()
    a = np.frombuffer(hdr, dtype=np.uint8)
    assert L.flacb200_read_streaminfo(C.c_void_p(a.ctypes.data), a.size, C.byref(si)) == 0
    assert (si.min_block_size, si.max_block_size, si.sample_rate, si.channels, si.bits_per_sample, si.total_samples) == (
        4096, 4096, 44100, 2, 16, total)
    assert si.n_seekpoints == 3 and si.frames_start == len(hdr) == 4 + 38 + 4 + 3 * 18 + 4 + 4096
    assert hdr[42] == 3 and hdr[42 + 4:42 + 4 + 8] == b"\xff" * 8     # placeholder points
    assert hdr[42 + 4 + 54] == 0x81                                    # PADDING is the last block
    # no GPU engine: the first write that completes a launch fails loudly, nothing is encoded on the CPU
    x = np.zeros(4096 * 2, dtype=np.int32)
    assert L.flacb200_writer_write_samples(h, C.c_void_p(x.ctypes.data), x.size) == 0      # buffered (launch_frames not reached)
    assert L.flacb200_writer_flush(h) == -1                                                  # FLACB200_E_NO_DEVICE
    L.flacb200_writer_close(h)
    # unknown total: no SEEKTABLE yet
    rc, h = _open(L, wo, 44100, 16, 2, 0)
    assert rc == 0
    assert L.flacb200_writer_header(h, C.byref(p), C.byref(n)) == 0
    assert n.value == 4 + 38 + 4 + 4096
    assert L.flacb200_writer_finalize(h) == 58    # NoSamples
    L.flacb200_writer_close(h)
