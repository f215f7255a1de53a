"""Device log / log2 (the glibc restatement in csrc/glibc_log.cuh, as the LPC kernels use it) against the C library on
the GPU box's host, bit for bit, over >= 1e8 inputs from the encoder's domain (src/encode.rs:3674, :3360)."""
import numpy as np
import pytest

from libm_domain import log2_inputs, log_inputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from flac_codec_b200 import Engine

    e = Engine(0)
    yield e
    e.close()


@pytest.mark.parametrize("fn,gen,total", [(0, log_inputs, 120_000_000), (1, log2_inputs, 100_000_000)])
def test_device_log_equals_glibc(eng, fn, gen, total):
    from oracle import oracle as fo

    chunk, done, seed = 20_000_000, 0, 100 * fn
    while done < total:
        x = gen(seed, chunk)
        want = fo.libm(fn, x)
        got = eng.debug_libm(fn, x)
        bad = np.flatnonzero(want.view(np.uint64) != got.view(np.uint64))
        assert bad.size == 0, (fn, seed, bad.size, x[bad[:5]].tolist(), want[bad[:5]].tolist(), got[bad[:5]].tolist())
        done += x.size
        seed += 1
