"""Host logic of the four drop-in classes, exercised WITHOUT a GPU: the ctypes classes are replaced
by oracle-backed fakes (tests/fake_backend.py), so these tests pin everything the host does around
the kernels -- the reference's random-stream order, NMC phase set-up, swap rules, output assembly,
side-effect files -- against the golden vectors of the unmodified reference."""
import os
import random

import numpy as np
import pytest

import fake_backend
from conftest import golden

EPS = np.finfo(float).eps


def seed_all(s):
    np.random.seed(s)
    random.seed(s)


@pytest.fixture
def fake_device(monkeypatch, tmp_cwd):
    fake_backend.install(monkeypatch)


def test_npt_sparse_input(fake_device):
    from nlmc_b200 import NPT
    from oracle import oracle as O
    g = golden("npt_run_c5")
    A, h = O.ea3d_pm_j(int(g["L"]), int(g["instance_seed"]))
    seed_all(int(g["seed"]))
    M, E = NPT(A, h).run(g["beta_list"], 6, [False] * 6, num_sweeps_MCMC=6, num_sweeps_read=6, num_swap_attempts=3,
                         num_swapping_pairs=2, num_cores=1)
    assert np.array_equal(M, g["M"].astype(float)) and np.array_equal(E, g["E"])


def test_npt_with_nmc_replicas(fake_device):
    from nlmc_b200 import NPT
    from oracle.make_golden import NPT_KW
    g = golden("npt_run_c2")
    seed_all(int(g["seed"]))
    obj = NPT(g["J"], g["h"])
    M, E = obj.run(g["beta_list"], 4, list(g["doNMC"]), num_sweeps_MCMC=60, num_sweeps_read=20, num_swap_attempts=4,
                   num_swapping_pairs=1, num_cores=1, **NPT_KW)
    assert np.array_equal(M, g["M"].astype(float)) and np.array_equal(E, g["E"])
    assert obj.num_sweeps_MCMC_per_swap == 15 and obj.num_sweeps_per_NMC_phase_per_swap == 3
    with pytest.raises(ValueError, match="length of doNMC"):
        NPT(g["J"], g["h"]).run(g["beta_list"], 4, [False] * 3)


@pytest.mark.parametrize("name", ["nmc_run_c1", "nmc_run_gauss"])
def test_nmc_run(fake_device, name):
    from nlmc_b200 import NMC
    g = golden(name)
    a = g["args"]
    seed_all(int(g["seed"]))
    M, E, mn = NMC(g["J"], g["h"]).run(int(a[0]), int(a[1]), int(a[2]), int(a[3]), int(a[4]), a[5], a[6], a[7], a[8],
                                       a[9], a[10], a[11], int(a[12]), a[13])
    assert np.array_equal(M, g["M"].astype(float))
    np.testing.assert_allclose(E, g["E"], rtol=1e-9)
    assert isinstance(mn, (float, np.float64)) and np.isclose(mn, float(g["min_energy"]), rtol=1e-9)


def test_apt_preprocessor(fake_device):
    from nlmc_b200 import APT_preprocessor
    g = golden("apt_preprocessor_c2")
    a = g["args"]
    seed_all(int(g["seed"]))
    beta, sigma = APT_preprocessor(g["J"].copy(), g["h"].copy()).run(int(a[0]), int(a[1]), int(a[2]), a[3], a[4], a[5],
                                                                     a[6], 0, 1)
    assert isinstance(beta, list) and isinstance(sigma, list)
    assert np.array_equal(np.array(beta, dtype=float), g["beta"]) and np.array_equal(np.array(sigma), g["sigma"])
    for f in ("beta_list_python.npy", "sigma_list_python.npy", "Results/data/Energy_iter_1.npy",
              "Results/data/sigma_iter_1.npy"):
        assert os.path.exists(f)
    assert np.array_equal(np.load("beta_list_python.npy"), g["beta"])
    with pytest.raises(ValueError):  # NPT/unittests/test_apt_preprocessor.py:45-50
        APT_preprocessor(g["J"].copy(), g["h"].copy()).run(num_sweeps_MCMC=-100)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_apt_icm(fake_device, tag):
    from nlmc_b200 import APT_ICM
    g = golden("apt_icm_c4")
    nsm, nsr, nsa, npairs = (int(v) for v in g[f"{tag}_args"])
    seed_all(int(g[f"{tag}_seed"]))
    obj = APT_ICM(g["J"].copy(), g["h"].copy())
    M, E = obj.run(g["beta_list"], 4, num_sweeps_MCMC=nsm, num_sweeps_read=nsr, num_swap_attempts=nsa,
                   num_swapping_pairs=npairs)
    assert obj.num_sweeps_MCMC == nsm and obj.h.shape == (64, 1)
    assert M.shape == (64 * 4, (nsm // nsa) * 10)
    assert np.array_equal(M, g[f"{tag}_M"].astype(float)) and np.array_equal(E, g[f"{tag}_E"])
    cl = obj.find_disagreement_clusters(g["s1"], g["s2"], g["J"])
    assert len(cl) == int(g["n_clusters"])
