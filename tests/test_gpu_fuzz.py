"""Randomised differential test (tools/fuzz_parity.py): over random stream shapes the GPU's frames are byte-identical to the
oracle's and both GPU decoders return the input PCM."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


@pytest.mark.parametrize("seed", [11, 12])
def test_random_streams_encode_and_decode_parity(seed):
    import fuzz_parity
    from flac_codec_b200 import Engine

    eng = Engine(0)
    try:
        assert fuzz_parity.run(120, seed, eng) is None
    finally:
        eng.close()
