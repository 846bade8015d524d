"""GPU: the device tanh / arctanh used by K5 (LBP) and K1 (replay) are bit-equal to np.tanh / np.arctanh on
> 9e6 arguments -- committed digests, explicit vectors, and live numpy when the host has the golden build's path."""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import npmath_vectors as V  # noqa: E402

pytestmark = pytest.mark.gpu

with open(os.path.join(HERE, "golden", "npmath_digests.json")) as f:
    GOLD = json.load(f)


def _fn(which):
    from nlmc_b200 import _lib
    return _lib.np_tanh if which == "tanh" else _lib.np_arctanh


def _same(a, b):
    return bool(np.all((a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))))


@pytest.mark.parametrize("which", ["tanh", "arctanh"])
def test_device_function_matches_committed_digests(which):
    sets = V.tanh_sets() if which == "tanh" else V.arctanh_sets()
    total = 0
    for name, x in sets.items():
        assert V.digest(_fn(which)(x)) == GOLD[which][name]["sha256"], (which, name)
        total += x.size
    assert total >= 4_000_000


@pytest.mark.parametrize("which", ["tanh", "arctanh"])
def test_device_function_matches_explicit_vectors(which):
    g = np.load(os.path.join(HERE, "golden", "npmath_vectors.npz"))
    x, y = g[which + "_x"].view(np.float64), g[which + "_y"].view(np.float64)
    assert _same(_fn(which)(x), y)


@pytest.mark.parametrize("which", ["tanh", "arctanh"])
def test_device_function_matches_live_numpy(which):
    if not V.numpy_is_golden_build():
        pytest.skip("this host's numpy does not take the AVX-512 path of the golden build")
    ref = np.tanh if which == "tanh" else np.arctanh
    sets = V.tanh_sets(1 << 18) if which == "tanh" else V.arctanh_sets(1 << 18)
    with np.errstate(all="ignore"):
        for name, x in sets.items():
            assert _same(_fn(which)(x), ref(x)), (which, name)


def test_atanh_saturated_public_method():
    from nlmc_b200.nmc import NMC
    J = np.array([[0.0, 1.0], [1.0, 0.0]])
    obj = NMC(J, np.zeros(2))
    x = np.array([-2.0, -1.0, -0.5, 0.0, 0.3, 1.0, 7.0])
    e = np.finfo(float).eps
    want = np.arctanh(np.clip(x, -1 + e, 1 - e))
    got = obj.atanh_saturated(x)
    if V.numpy_is_golden_build():
        assert _same(np.asarray(got), want)
    else:
        np.testing.assert_allclose(got, want, rtol=1e-15)
    assert np.ndim(obj.atanh_saturated(0.25)) == 0
