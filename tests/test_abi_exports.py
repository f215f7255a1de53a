"""CPU-side checks of the drop-in boundary: libflacb200.so loads, exports every symbol that
include/flacb200.h declares, the option presets mirror Options::{default,fast,best}
(src/encode.rs:1376-1408, :1635-1657), and -- without a GPU -- every compute entry point fails
loudly instead of falling back to a CPU path."""
import ctypes as C
import os
import re

import pytest

from flacb200_testutil import ROOT


def declared_symbols():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if not fn.endswith(".h"):
            continue
        with open(os.path.join(ROOT, "include", fn)) as f:
            text = f.read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names.update(re.findall(r"\b(flacb200_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


@pytest.fixture(scope="module")
def lib():
    from flac_codec_b200 import build

    build.build()
    from flac_codec_b200 import _abi

    return _abi.lib()


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert len(names) >= 20
    raw = C.CDLL(os.path.join(ROOT, "flac_codec_b200", "libflacb200.so"))
    missing = [n for n in names if not hasattr(raw, n)]
    assert not missing, missing


def test_binding_lists_every_declared_symbol(lib):
    from flac_codec_b200 import _abi

    assert sorted(_abi.EXPORTS) == declared_symbols()


def test_option_presets_mirror_reference(lib):
    from flac_codec_b200 import Options

    d, f, b = Options.default().c, Options.fast().c, Options.best().c
    assert (d.block_size, d.max_lpc_order, d.max_partition_order, d.mid_side, d.exhaustive_channel_correlation,
            d.window_kind, d.tukey_p) == (4096, 8, 5, 1, 1, 2, 0.5)
    assert (f.block_size, f.max_lpc_order, f.max_partition_order, f.mid_side, f.exhaustive_channel_correlation) == (
        1152, 0, 3, 0, 0)
    assert (b.block_size, b.max_lpc_order, b.max_partition_order, b.mid_side, b.exhaustive_channel_correlation) == (
        4096, 12, 6, 1, 1)
    with pytest.raises(ValueError):
        Options.default().block_size(15)       # OptionsError::InvalidBlockSize  (:1418)
    with pytest.raises(ValueError):
        Options.default().max_lpc_order(33)    # OptionsError::InvalidLpcOrder   (:1430)
    with pytest.raises(ValueError):
        Options.default().max_partition_order(16)


def test_strerror_names_follow_error_enum(lib):
    assert lib.flacb200_strerror(0) == b"Ok"
    assert lib.flacb200_strerror(39) == b"Crc8Mismatch"
    assert lib.flacb200_strerror(40) == b"Crc16Mismatch"
    assert lib.flacb200_strerror(21) == b"ShortBlock"
    assert lib.flacb200_strerror(60) == b"ResidualOverflow"
    assert b"no CUDA device" in lib.flacb200_strerror(-1)


def test_no_cpu_fallback_without_gpu(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from flac_codec_b200 import Engine, _abi

    with pytest.raises(_abi.FlacB200Error) as ei:
        Engine(0)
    assert ei.value.code == -1   # FLACB200_E_NO_DEVICE


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "flac_codec_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                with open(os.path.join(dirpath, fn)) as f:
                    text = f.read()
                assert "flac_oracle" not in text and "from oracle" not in text and "import oracle" not in text, fn
