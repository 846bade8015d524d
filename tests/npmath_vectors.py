"""Argument sets for the np.tanh / np.arctanh parity tests (CPU and GPU share them).

Arguments come from a splitmix64 stream written out in numpy integer arithmetic, so they are identical
everywhere; the expected outputs are pinned by tests/golden/npmath_digests.json (sha256 of the output bit
patterns per set, generated with numpy 2.3.5 on an AVX-512 host by tests/make_npmath_golden.py) and by the
explicit vectors in tests/golden/npmath_vectors.npz.
"""
from __future__ import annotations

import hashlib

import numpy as np

N_PER_SET = 1 << 20


def splitmix64(seed: int, n: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) + np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _unit(bits):  # uniform in [0, 1) with 53 bits
    return (bits >> np.uint64(11)).astype(np.float64) * 2.0 ** -53


def tanh_sets(n: int = N_PER_SET) -> dict:
    a, b, c = splitmix64(11, n), splitmix64(12, n), splitmix64(13, n)
    sets = {
        "uniform_pm40": (_unit(a) * 80.0 - 40.0),
        "uniform_pm3": (_unit(b) * 6.0 - 3.0),
        "raw_bits": c.view(np.float64).copy(),                                   # every exponent, NaNs, infinities
        "log_scale": (_unit(a) * 2.0 - 1.0) * 2.0 ** ((b % np.uint64(1100)).astype(np.float64) - 1070.0),
        "interval_edges": None,
    }
    # the 16 interval boundaries of the routine (exponent + first mantissa bit) and their neighbours
    edges = []
    for e in range(0x3fa, 0x406):
        for top in (0, 1):
            base = (e << 52) | (top << 51)
            for d in (-2, -1, 0, 1, 2):
                edges += [base + d, (base + d) | (1 << 63)]
    edges += [0, 1 << 63, 1, 0x7ff0000000000000, 0xfff0000000000000, 0x7ff8000000000000, 0x7fefffffffffffff,
              0x7fe0000000000000, 0x7fe8000000000000]
    sets["interval_edges"] = np.array(edges, dtype=np.uint64).view(np.float64)
    return sets


def arctanh_sets(n: int = N_PER_SET) -> dict:
    a, b, c = splitmix64(21, n), splitmix64(22, n), splitmix64(23, n)
    sign = np.where((c & np.uint64(1)) == 0, 1.0, -1.0)
    raw = a.view(np.float64).copy()
    raw = np.where(np.abs(raw) < 1.0, raw, np.ldexp(_unit(b), -(c % np.uint64(60)).astype(np.int64)))
    sets = {
        "uniform_pm1": _unit(a) * 2.0 - 1.0,
        "near_one": sign * (1.0 - np.ldexp(_unit(a) + 2.0 ** -53, -(b % np.uint64(53)).astype(np.int64))),
        "raw_bits_below_one": raw,
        "log_scale": sign * _unit(a) * 2.0 ** (-(b % np.uint64(1000)).astype(np.float64)),
        # the products the LBP update feeds to arctanh (nmc.py:205)
        "tanh_products": np.tanh(_unit(a) * 6.0 - 3.0) * np.tanh(_unit(b) * 40.0 - 20.0),
        "reciprocal_steps": None,
    }
    steps = [0x040f0, 0x0c980, 0x15b40, 0x1f700, 0x29e60, 0x35240, 0x41430, 0x4e600,
             0x5c990, 0x6c160, 0x7d070, 0x8f9d0, 0xa41a0, 0xbad10, 0xd41c0, 0xf0820]
    xs = []
    for s in steps:
        for d in range(-3, 4):
            for lo in (0, 1, 0x80000000, 0xffffffff):
                y = np.array([(0x3ff << 52) | ((s + d) << 32) | lo], dtype=np.uint64).view(np.float64)[0]
                xs += [y - 1.0, 1.0 - y / 2, 1.0 - y / 4, 1.0 - y / 1024, 1.0 - y * 2.0 ** -40]
    xs = np.array(xs + [0.0, 5e-324, 2.0 ** -1022, 1.0 - 2.0 ** -53, 1.0 - 2.0 ** -52, 0.5, 2.0 ** -30])
    sets["reciprocal_steps"] = np.concatenate([xs, -xs])
    return sets


def digest(out: np.ndarray) -> str:
    """sha256 of the output bit patterns, every NaN mapped to one pattern (payloads are not part of the contract)."""
    o = np.ascontiguousarray(out, dtype=np.float64).copy()
    o[np.isnan(o)] = np.nan
    return hashlib.sha256(o.view(np.uint64).tobytes()).hexdigest()


def numpy_is_golden_build() -> bool:
    """True when this host's numpy takes the code path the digests were made with (AVX-512 dispatch)."""
    import os
    from json import load
    x = np.array([0.7310585786300049, -0.33, 0.9999999, 1e-5, 0.123456789])
    with open(os.path.join(os.path.dirname(__file__), "golden", "npmath_digests.json")) as f:
        g = load(f)
    with np.errstate(all="ignore"):
        return (np.tanh(x).view(np.uint64).tolist() == g["canary"]["tanh"]
                and np.arctanh(x).view(np.uint64).tolist() == g["canary"]["arctanh"])
