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


# ------------------------------------------------------------------ element-level public methods (path_methods.py)
def test_public_mcmc_and_tasks(fake_device):
    from nlmc_b200 import NMC, NPT, APT_ICM, APT_preprocessor
    e = golden("mcmc_element")
    for tag in ("pm_anneal", "gauss_fixed"):
        J, h = e[f"{tag}_J"], e[f"{tag}_h"]
        seed_all(int(e[f"{tag}_seed"]))
        m0 = np.sign(2 * np.random.rand(len(h)) - 1)
        M = NMC(J, h).MCMC(int(e[f"{tag}_sweeps"]), m0, float(e[f"{tag}_beta"]), J, h, anneal=bool(e[f"{tag}_anneal"]))
        assert np.array_equal(M, e[f"{tag}_M"])
    g = golden("public_methods")
    J, h, ms = g["J"], g["h"], g["lbp_m_star"].astype(float)
    npt = NPT(J, h)
    seed_all(int(g["task_seed"]))
    Mt = npt.MCMC_task(2, 6, ms.copy(), g["task_betas"])
    assert np.array_equal(Mt, g["task_M"])
    mn, EE1 = npt.replica_energy(Mt, 4)
    np.testing.assert_allclose(EE1, g["rep_EE1"], rtol=1e-12)
    a = g["nmctask_args"]
    seed_all(int(g["nmctask_seed"]))
    Mn = npt.NMC_task(ms.copy(), int(a[0]), int(a[1]), int(a[2]), int(a[3]), a[4], a[5], a[6], a[7], a[8], a[9], a[10],
                      int(a[11]), a[12])
    assert np.array_equal(Mn, g["nmctask_M"])
    prep, icm = APT_preprocessor(J, h), APT_ICM(J, h)
    seed_all(int(g["apt_seed"]))
    Mp = prep.MCMC(5, ms.copy(), 1.2)
    En, mlast = prep.MCMC_task(ms.copy(), 0.8, 7, 3)
    Mi = icm.MCMC(4, ms.copy(), 0.6)
    assert np.array_equal(Mp, g["prep_M"]) and np.array_equal(Mi, g["icm_M"]) and np.array_equal(mlast, g["prep_task_m"])
    np.testing.assert_allclose(En, g["prep_task_E"], rtol=1e-12)
    np.testing.assert_allclose(icm.replica_energy(Mi, 4)[1], g["icm_rep_EE1"], rtol=1e-12)
    with pytest.raises(ValueError, match="LRUCache"):
        npt.MCMC(2, ms, 1.0, J, h, hash_table=object(), use_hash_table=True)


@pytest.mark.parametrize("tag,field,init", [("lbp", "lbp_field1", "cold"), ("lbp2", "lbp_field2", "warm"),
                                            ("lbp3", None, "dense")])
def test_public_loopy_belief_propagation(fake_device, tag, field, init):
    """The dense <-> edge conversion around K5, including messages off the entries of J (explicit-zero pattern)."""
    from nlmc_b200 import NMC
    g = golden("public_methods")
    J, ms = g["J"], g["lbp_m_star"].astype(float)
    n = len(ms)
    if init == "cold":
        h0, u0, hf, it_max = np.zeros((n, n)), J * ms.reshape(1, -1), g[field], int(g["lbp_max_iter"])
    elif init == "warm":
        h0, u0, hf, it_max = g["lbp_h_msgs"], g["lbp_u_msgs"], g[field], int(g["lbp_max_iter"])
    else:
        h0, u0, hf, it_max = g["lbp3_h0"], g["lbp3_u0"], g["h"], 3
    out = NMC(J, g["h"]).LoopyBeliefPropagation(J, hf, 1.5, h0.copy(), u0.copy(), 1e-10, it_max)
    assert out[4] == int(g[f"{tag}_iteration"])
    for got, key in zip(out, ("marg", "corr", "h_tilde", "J_tilde", None, "h_msgs", "u_msgs")):
        if key:
            np.testing.assert_allclose(got, g[f"{tag}_{key}"], rtol=1e-12, atol=1e-13, err_msg=key)


def test_public_lbp_convexified_and_subroutine(fake_device):
    from nlmc_b200 import NMC, NPT
    g = golden("public_methods")
    J, h, ms = g["J"], g["h"], g["lbp_m_star"].astype(float)
    a = g["conv_args"]
    epsv = np.abs(h) + np.sum(np.abs(J), axis=1)
    cl, marg, mean, ht, jt = NMC(J, h).LBP_convexified(a[0], a[1], a[2], ms.copy(), epsv, a[3], int(a[4]), a[5], a[6], a[7])
    lams = list(marg.keys())
    assert np.array_equal(np.array(lams), g["conv_lambdas"])
    assert np.array_equal(np.array([marg[k] for k in lams]), g["conv_marginals"])
    np.testing.assert_allclose(np.array([ht[k] for k in lams]), g["conv_h_tilde"], rtol=1e-12)
    np.testing.assert_allclose(jt[lams[-1]], g["conv_J_tilde_last"], rtol=1e-12, atol=1e-14)
    assert np.array_equal(np.concatenate(cl), g["conv_clusters_flat"])
    s = g["sub_args"]
    for variant, cls in (("nmc", NMC), ("npt", NPT)):
        seed_all(int(g["sub_seed"]))
        M, E, mn, ac = cls(J, h).NMC_subroutine(ms.copy(), int(s[0]), int(s[1]), int(s[2]), int(s[3]), s[4], s[5], s[6],
                                                s[7], s[8], s[9], s[10], int(s[11]), s[12],
                                                all_clusters=g["sub_clusters"].copy())
        assert np.array_equal(M, g[f"sub_{variant}_M"]) and np.array_equal(ac, g[f"sub_{variant}_clusters"])
        np.testing.assert_allclose(E, g[f"sub_{variant}_E"], rtol=1e-12)


# ------------------------------------------------------------------ production-mode host logic (production.py)
@pytest.fixture
def fake_engines(fake_device, monkeypatch):
    from nlmc_b200 import production
    monkeypatch.setattr(production, "_generic_engine", lambda prob, betas, seed: fake_backend.FakeEngine(prob, betas, seed))
    monkeypatch.setattr(production, "_msc_eligible", lambda prob: False)  # everything through the generic host path


def test_production_host_logic_npt_and_nmc(fake_engines):
    """NPT.run / NMC.run in production mode on a stand-in engine: return contracts, energies that belong to the returned
    states, NMC replicas at global_beta, and only the last round recorded."""
    from nlmc_b200 import NMC, NPT
    from oracle import oracle as O
    J, h = O.random_pm_graph(24, 0.25, 3)
    h = 0.2 * np.random.RandomState(1).randn(24)
    csr = O.Csr(J)
    seed_all(4)
    betas = np.array([0.4, 0.8, 1.2, 1.6])
    M, E = NPT(J, h, mode="production").run(betas, 4, [False, False, True, True], num_sweeps_MCMC=24, num_sweeps_read=12,
                                            num_swap_attempts=3, num_swapping_pairs=1, num_cycles=2, global_beta=2.0,
                                            lambda_start=3, threshold_initial=0.99, threshold_cutoff=0.9,
                                            max_iterations=50, tolerance=1e-9)
    n, spm = 24, 8
    assert M.shape == (4 * n, spm) and E.shape == (4,) and np.all(np.abs(M) == 1)
    for r in range(4):
        Er = O.energy(csr, h, M[r * n:(r + 1) * n, :4].T.astype(np.int8))   # num_sweeps_read // num_swap_attempts = 4
        assert E[r] == pytest.approx(Er.min(), rel=1e-9)
    seed_all(5)
    Mo, Eo, mn = NMC(J, h, mode="production").run(10, 6, 2, 1, 2, 10.0, 2.0, 3, 0.01, 0.9, 0.99, 0.9, 50, 1e-9)
    assert Mo.shape == (n, 2 * 3 * 3) and Eo.shape == (18,) and mn == Eo.min()
    np.testing.assert_allclose(Eo, O.energy(csr, h, Mo.T.astype(np.int8)), rtol=1e-9)


def test_production_host_logic_preprocessor_and_icm(fake_engines):
    from nlmc_b200 import APT_ICM, APT_preprocessor
    from oracle import oracle as O
    J, h = O.random_pm_graph(20, 0.3, 7)
    seed_all(6)
    beta, sigma = APT_preprocessor(J, h, mode="production").run(num_sweeps_MCMC=12, num_sweeps_read=8, num_rng=5,
                                                               beta_start=0.5, alpha=1.25, sigma_E_val=1000, beta_max=2.0,
                                                               use_hash_table=0, num_cores=1)
    assert len(beta) >= 2 and np.all(np.diff(beta) > 0) and len(sigma) in (len(beta), len(beta) - 1)
    assert os.path.exists("beta_list_python.npy") and os.path.exists(os.path.join("Results", "data", "Energy_iter_1.npy"))
    assert np.load(os.path.join("Results", "data", "Energy_iter_1.npy")).shape == (5, 8)
    seed_all(7)
    for spm in (1, 2):
        M, E = APT_ICM(J, h, mode="production").run(np.array([0.4, 0.9, 1.4]), 3, num_sweeps_MCMC=3 * spm,
                                                   num_sweeps_read=3 * spm, num_swap_attempts=3, num_swapping_pairs=1)
        assert M.shape == (3 * 20, spm * 10) and E.shape == (3,) and np.all(np.abs(M) == 1)
        csr = O.Csr(J)
        for r in range(3):
            Er = O.energy(csr, np.zeros(20), M[r * 20:(r + 1) * 20, :spm].T.astype(np.int8))
            assert E[r] == pytest.approx(Er.min(), rel=1e-9)


def test_production_host_logic_bit_packed_paths(fake_device, monkeypatch):
    """The host side of the bit-packed production paths (NPT.run without NMC replicas incl. num_runs, APT_preprocessor,
    APT_ICM on a lattice) on a stand-in for _lib.Msc."""
    from nlmc_b200 import APT_ICM, APT_preprocessor, NPT, _lib
    from oracle import oracle as O
    monkeypatch.setattr(_lib, "Msc", fake_backend.FakeMsc)
    A, h = O.ea3d_pm_j(3, 2)
    n = 27
    csr = O.Csr(A)
    betas = np.array([0.4, 0.9, 1.4])
    seed_all(8)
    obj = NPT(A, h, mode="production")
    obj.num_runs = 2
    M, E = obj.run(betas, 3, [False] * 3, num_sweeps_MCMC=12, num_sweeps_read=6, num_swap_attempts=3, num_swapping_pairs=1)
    assert M.shape == (3 * n, 4) and E.shape == (3,) and obj.energies_all_runs.shape == (3, 2)
    for r in range(3):
        assert E[r] == O.energy(csr, h, M[r * n:(r + 1) * n, :2].T.astype(np.int8)).min()
    seed_all(9)
    beta, sigma = APT_preprocessor(A, h, mode="production").run(num_sweeps_MCMC=8, num_sweeps_read=6, num_rng=3, beta_start=0.5,
                                                               alpha=1.25, sigma_E_val=1000, beta_max=1.5, use_hash_table=0)
    assert len(beta) >= 2 and np.all(np.diff(beta) > 0)
    seed_all(10)
    M, E = APT_ICM(A.toarray(), h, mode="production").run(betas, 3, num_sweeps_MCMC=6, num_sweeps_read=6, num_swap_attempts=3,
                                                         num_swapping_pairs=1)
    assert M.shape == (3 * n, 2 * 10) and np.all(np.abs(M) == 1)
    for r in range(3):
        assert E[r] == O.energy(csr, np.zeros(n), M[r * n:(r + 1) * n, :2].T.astype(np.int8)).min()


def test_production_host_logic_hybrid_ladder(fake_device, monkeypatch):
    """NPT.run on a +-J lattice with NMC on the coldest replicas (C2-shaped): plain replicas on the bit-packed stand-in,
    NMC replicas on the generic stand-in, exchanges between the two kinds -- return contract and energies that belong to
    the returned states."""
    from nlmc_b200 import NPT, _lib, production
    from oracle import oracle as O
    monkeypatch.setattr(_lib, "Msc", fake_backend.FakeMsc)
    monkeypatch.setattr(production, "_generic_engine", lambda prob, betas, seed: fake_backend.FakeEngine(prob, betas, seed))
    A, h = O.ea3d_pm_j(3, 4)
    n = 27
    csr = O.Csr(A)
    betas = np.array([0.3, 0.7, 1.1, 1.5])
    seed_all(12)
    obj = NPT(A.toarray(), h, mode="production")
    M, E = obj.run(betas, 4, [False, False, True, True], num_sweeps_MCMC=36, num_sweeps_read=18, num_swap_attempts=3,
                   num_swapping_pairs=1, num_cycles=2, global_beta=2.0, lambda_start=3, threshold_initial=0.99,
                   threshold_cutoff=0.9, max_iterations=50, tolerance=1e-9)
    spm, spr = 12, 6
    assert M.shape == (4 * n, spm) and E.shape == (4,) and np.all(np.abs(M) == 1)
    for r in range(4):
        Er = O.energy(csr, h, M[r * n:(r + 1) * n].T.astype(np.int8))
        assert E[r] == Er[:spr].min()
