"""Inputs for the log / log2 parity tests: the encoder's domain and the places where a table-based libm is fragile."""
import numpy as np


def log_inputs(seed: int, n: int) -> np.ndarray:
    """Arguments of f64::ln in subframe_bits_by_order (src/encode.rs:3674): error * 0.5 / n, any positive double in
    practice; drawn log-uniform over [1e-12, 1e12], plus the neighbourhood of 1.0 (the function's
    separate branch), table-interval edges and exact powers of two."""
    rng = np.random.default_rng(seed)
    k = n // 4
    a = np.exp(rng.uniform(np.log(1e-12), np.log(1e12), n - 3 * k))
    b = 1.0 + rng.uniform(-0.07, 0.07, k)                                      # the close-to-1 polynomial and its borders
    # mantissas next to the 128 (log) / 64 (log2) table boundaries, random exponents
    m = (rng.integers(0, 128, k) / 128.0 + 1.0) * (1.0 + rng.integers(-40, 41, k) * 2.0 ** -52)
    c = np.ldexp(m, rng.integers(-40, 41, k))
    d = np.ldexp(1.0 + rng.integers(-3, 4, k) * 2.0 ** -52, rng.integers(-60, 61, k))   # powers of two +- a few ulp
    return np.concatenate([a, b, c, np.abs(d)])


def log2_inputs(seed: int, n: int) -> np.ndarray:
    """Arguments of f64::log2 in quantize (src/encode.rs:3360): the largest |LPC coefficient|; floor() of the result is
    what matters, so values within a few ulp of powers of two carry the weight."""
    rng = np.random.default_rng(seed + 1)
    k = n // 4
    a = np.exp(rng.uniform(np.log(1e-9), np.log(1e6), n - 3 * k))
    b = 1.0 + rng.uniform(-0.05, 0.05, k)
    m = (rng.integers(0, 64, k) / 64.0 + 1.0) * (1.0 + rng.integers(-40, 41, k) * 2.0 ** -52)
    c = np.ldexp(m, rng.integers(-30, 21, k))
    d = np.ldexp(1.0 + rng.integers(-8, 9, k) * 2.0 ** -52, rng.integers(-30, 21, k))
    return np.concatenate([a, b, c, np.abs(d)])
