"""flac_codec_b200 -- B200-native FLAC frame engine behind flac-codec's encode/decode API.

Python here is harness-level plumbing over the C ABI (include/flacb200.h); the product is
libflacb200.so (hand-written sm_100a CUDA).  Importing this package never imports the oracle.
"""
from . import _abi  # noqa: F401
from .engine import Engine, Options  # noqa: F401

__all__ = ["Engine", "Options", "_abi"]
