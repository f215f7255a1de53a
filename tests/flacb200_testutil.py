"""Shared helpers for the test-suite (fixtures, deterministic signal generators)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_DATA = os.path.join(GOLDEN, "ref_data")


def ref_file(name: str) -> bytes:
    with open(os.path.join(REF_DATA, name), "rb") as f:
        return f.read()
