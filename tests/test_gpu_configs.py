"""BASELINE.json configs C1, C2, C3 and C5 at their stated sizes on the GPU, against the oracle (tests/config_legs.py);
C4 is bench.py's workload.  The same legs are reported in bench.py's JSON line as `configs`."""
import pytest

import config_legs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from flac_codec_b200 import Engine

    e = Engine(0)
    e.set_keep_info(False)
    yield e
    e.close()


@pytest.fixture(scope="module")
def fo():
    from oracle import oracle

    return oracle


def _check(legs, n):
    assert len(legs) == n
    for leg in legs:
        assert leg["identical"] is True, leg
        assert leg["msamples_per_s"] > 0


def test_c1_whole_file_and_c2_decode_at_60_s(eng, fo):
    legs = config_legs.c1_c2(eng, fo, reps=1)
    _check(legs, 2)
    assert legs[0]["frames"] == 646 and legs[0]["md5_ok"]   # 645 x 4096 + 4080


def test_c3_all_1407_frames(eng, fo):
    legs = config_legs.c3(eng, fo, reps=1)
    _check(legs, 1)
    assert legs[0]["frames"] == 1407 and legs[0]["equal_frame_sizes"] == 1407 and legs[0]["size_delta"] == 0


def test_c5_order_32_streams_and_fixtures(eng, fo):
    legs = config_legs.c5(eng, fo, reps=1)
    _check(legs, 3 * len(config_legs.C5_SHAPES) + 4)
