"""The lower bound that lets k_analyze3 skip the exact size of the fixed predictor (flac_codec_b200/csrc/rice.cuh,
aw_choose_partitions_flat: lb_out): for a partition of n residuals with Rice parameter k and S = sum |r|,

    exact = sum ((zigzag(r) >> k) + 1 + k)   >=   max(0, ceil(2 S / 2^k) - n) + n (k + 1)

because zigzag(r) is 2|r| or 2|r| - 1 and floor(u / 2^k) >= u / 2^k - 1 + 2^-k.  The kernel compares the EXACT LPC size with this
bound of the fixed size; `lpc_bits < fixed_bits` (src/encode.rs:2931) is then decided without the exact fixed size.  Checked
here on the host, over random partitions, together with how tight the bound is (never more than n bits below the exact size)."""
import numpy as np


def zigzag(r):
    r = r.astype(np.int64)
    return np.where(r >= 0, 2 * r, -2 * r - 1)


def exact_bits(r, k):
    return int(((zigzag(r) >> k) + 1 + k).sum())


def lower_bound(r, k):
    n, s = r.size, int(np.abs(r.astype(np.int64)).sum())
    t = (2 * s + (1 << k) - 1) >> k
    return max(0, t - n) + n * (k + 1)


def test_lower_bound_holds_and_is_within_one_bit_per_sample():
    rng = np.random.default_rng(20261018)
    for trial in range(4000):
        n = int(rng.integers(1, 300))
        kind = trial % 5
        if kind == 0:
            r = rng.integers(-(1 << 24), 1 << 24, n)
        elif kind == 1:
            r = np.rint(rng.laplace(0, float(1 << int(rng.integers(0, 20))), n)).astype(np.int64)
        elif kind == 2:
            r = -np.abs(rng.integers(0, 1 << int(rng.integers(1, 25)), n))   # all negative: zigzag is 2|r| - 1 everywhere
        elif kind == 3:
            r = np.zeros(n, dtype=np.int64)
            r[rng.integers(0, n)] = int(rng.integers(-(1 << 30), 1 << 30))   # one outlier
        else:
            r = rng.integers(-3, 4, n)
        for k in (0, 1, 2, int(rng.integers(0, 31)), 30):
            e, lb = exact_bits(r, k), lower_bound(r, k)
            assert lb <= e, (trial, k)
            assert e - lb <= n, (trial, k)


def test_bound_is_attained():
    # all residuals negative odd multiples: every code loses the full 1 - 2^-k to the floor
    r = np.full(64, -(1 << 10))          # zigzag = 2^11 - 1
    k = 11
    assert exact_bits(r, k) == 64 * (0 + 1 + k)
    assert lower_bound(r, k) == 64 * (k + 1)
