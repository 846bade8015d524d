"""CPU: the restated np.tanh / np.arctanh (csrc/nlmc_npmath.h, host build in oracle/npmath_host.c) are bit-equal to
numpy on > 9e6 arguments -- against the committed digests everywhere, and against live numpy when this host's
numpy takes the AVX-512 code path the goldens were made with."""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import npmath_vectors as V  # noqa: E402
from oracle import oracle as O  # noqa: E402

with open(os.path.join(HERE, "golden", "npmath_digests.json")) as f:
    GOLD = json.load(f)


def _same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return bool(np.all((a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))))


@pytest.mark.parametrize("which", ["tanh", "arctanh"])
def test_restatement_matches_committed_digests(which):
    sets = V.tanh_sets() if which == "tanh" else V.arctanh_sets()
    assert set(sets) == set(GOLD[which])
    total = 0
    for name, x in sets.items():
        assert x.size == GOLD[which][name]["n"]
        assert V.digest(O.npmath_host(which, x)) == GOLD[which][name]["sha256"], (which, name)
        total += x.size
    assert total >= 4_000_000


@pytest.mark.parametrize("which", ["tanh", "arctanh"])
def test_restatement_matches_explicit_vectors(which):
    g = np.load(os.path.join(HERE, "golden", "npmath_vectors.npz"))
    x, y = g[which + "_x"].view(np.float64), g[which + "_y"].view(np.float64)
    assert x.size >= 8000
    assert _same(O.npmath_host(which, x), y)


@pytest.mark.parametrize("which", ["tanh", "arctanh"])
def test_restatement_matches_live_numpy(which):
    if not V.numpy_is_golden_build():
        pytest.skip("this host's numpy does not take the AVX-512 path of the golden build")
    fn = np.tanh if which == "tanh" else np.arctanh
    sets = V.tanh_sets(1 << 18) if which == "tanh" else V.arctanh_sets(1 << 18)
    with np.errstate(all="ignore"):
        for name, x in sets.items():
            assert _same(O.npmath_host(which, x), fn(x)), (which, name)
        # the arguments LBP really produces: tanh(beta J) tanh(beta h) at the README betas
        rng = np.random.RandomState(5)
        for beta in (2.5, 3.0, 13.6):
            t = np.tanh(beta * rng.choice([-1.0, 1.0], 100000)) * np.tanh(beta * rng.randn(100000) * 3)
            arg = np.clip(t, -1 + 2.0 ** -52, 1 - 2.0 ** -52) if which == "arctanh" else beta * rng.randn(100000) * 3
            assert _same(O.npmath_host(which, arg), fn(arg))
