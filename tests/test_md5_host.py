"""flacb200_md5_many (csrc/md5_mb.cpp: eight streams per AVX2 register, host threads) against hashlib -- host code, no GPU."""
import hashlib

import numpy as np


def test_md5_many_equals_hashlib():
    import ctypes as C

    from flac_codec_b200 import _abi

    rng = np.random.default_rng(5)
    bufs = [rng.integers(0, 256, int(n), dtype=np.uint8) for n in (0, 1, 55, 56, 63, 64, 65, 1000, 4096 * 64 + 3, 100000, 100001, 7, 999999, 64 * 300, 12, 130, 131)]
    ptrs = (C.c_void_p * len(bufs))(*[b.ctypes.data for b in bufs])
    lens = (C.c_size_t * len(bufs))(*[b.size for b in bufs])
    for threads in (1, 3, 0):
        out = np.zeros(16 * len(bufs), dtype=np.uint8)
        _abi.lib().flacb200_md5_many(ptrs, lens, len(bufs), C.c_void_p(out.ctypes.data), threads)
        for i, b in enumerate(bufs):
            assert out[16 * i:16 * i + 16].tobytes() == hashlib.md5(b.tobytes()).digest(), (threads, i)
