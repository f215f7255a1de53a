"""The restatement of glibc's log / log2 (flac_codec_b200/csrc/glibc_log.cuh) against the C library itself, on the CPU:
the header is compiled for the host (plain IEEE operations + libm's exact fma) and must agree bit for bit.  The device
build of the same header is compared on the GPU in tests/test_gpu_libm.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from libm_domain import log2_inputs, log_inputs

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def port(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("glog") / "glibc_log_host.so")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", so,
                           os.path.join(HERE, "glibc_log_host.cpp"), "-lm"])
    L = C.CDLL(so)
    L.port_libm.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
    L.port_libm.restype = None

    def run(fn, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty_like(x)
        L.port_libm(fn, x.ctypes.data, out.ctypes.data, x.size)
        return out
    return run


@pytest.mark.parametrize("fn,gen", [(0, log_inputs), (1, log2_inputs)])
def test_port_equals_glibc_bit_for_bit(port, fn, gen):
    from oracle import oracle as fo

    for seed in range(2):
        x = gen(seed, 4_000_000)
        want = fo.libm(fn, x)
        got = port(fn, x)
        bad = np.flatnonzero(want.view(np.uint64) != got.view(np.uint64))
        assert bad.size == 0, (fn, bad.size, x[bad[:5]].tolist(), want[bad[:5]].tolist(), got[bad[:5]].tolist())


def test_port_special_values(port):
    from oracle import oracle as fo

    x = np.array([1.0, 2.0, 0.5, 4.0, 1e-300, 5e-324, 2.2250738585072014e-308, 1.7976931348623157e308, np.inf, 0.0,
                  0.9375, 1.0646972656249998, 1.0646972656250000, 0.93749999999999989], dtype=np.float64)
    for fn in (0, 1):
        assert np.array_equal(fo.libm(fn, x).view(np.uint64), port(fn, x).view(np.uint64)), fn
