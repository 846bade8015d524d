"""Pins the CPU oracle (oracle/) against the golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py).  Runs everywhere -- no GPU, no /root/reference needed."""
import itertools
import random

import numpy as np
import pytest

from conftest import golden
from oracle import oracle as O

EPS = np.finfo(float).eps


def seed_all(s):
    np.random.seed(s)
    random.seed(s)


@pytest.mark.parametrize("tag", ["pm_fixed", "pm_anneal", "gauss_fixed", "gauss_anneal"])
def test_mcmc_element(tag):
    g = golden("mcmc_element")
    J, h = g[f"{tag}_J"], g[f"{tag}_h"]
    seed_all(int(g[f"{tag}_seed"]))
    m0 = np.sign(2 * np.random.rand(len(h)) - 1)
    sched = O.anneal_schedule(int(g[f"{tag}_sweeps"]), float(g[f"{tag}_beta"]), bool(g[f"{tag}_anneal"]), 1, 0)
    M, _ = O.mcmc(O.Csr(J), h, m0, sched)
    assert np.array_equal(M.T, g[f"{tag}_M"])
    E = O.energy(O.Csr(J), h, M)
    if tag.startswith("pm"):
        assert np.array_equal(E, g[f"{tag}_E"])  # integer energies: bit-exact
    else:
        np.testing.assert_allclose(E, g[f"{tag}_E"], rtol=1e-9)  # north_star tolerance for Gaussian J


def test_lbp_marginals_and_clusters():
    g = golden("lbp")
    J, h, ms = g["J"], g["h"], g["m_star"].astype(float)
    csr = O.Csr(J)
    epsv = np.abs(h) + np.sum(np.abs(J), axis=1)
    assert np.array_equal(O._pairwise_rowsum_abs(csr), np.sum(np.abs(J), axis=1))
    lam0, lam_end, fac, tol, max_it, thr_i, thr_c = g["params"]
    beta = float(g["beta"])
    # every lambda step: identical marginal and identical iteration count
    u = np.ascontiguousarray(csr.val * ms[csr.ci])
    hm, tot = np.zeros_like(u), np.zeros(csr.n)
    for lam, marg_ref, it_ref in zip(g["lambdas"], g["marginals"], g["iters"]):
        marg, it = O.lbp(csr, np.ascontiguousarray(h + lam * ms * epsv), beta, u, hm, tot, tol, int(max_it))
        assert it == it_ref
        if it != int(max_it) - 1:  # on divergence the reference records the previous marginal
            assert np.array_equal(marg, marg_ref)
    cl, _, steps = O.lbp_convexified(csr, h, ms, epsv, lam0, lam_end, fac, tol, int(max_it), thr_i, thr_c, beta)
    assert steps == len(g["lambdas"])
    assert np.array_equal(np.concatenate(cl), g["clusters"])


@pytest.mark.parametrize("name", ["nmc_run_c1", "nmc_run_gauss"])
def test_nmc_run(name):
    g = golden(name)
    a = g["args"]
    seed_all(int(g["seed"]))
    M, E, mn = O.nmc_run(g["J"], g["h"], int(a[0]), int(a[1]), int(a[2]), int(a[3]), int(a[4]), a[5], a[6], a[7],
                         a[8], a[9], a[10], a[11], int(a[12]), a[13])
    assert np.array_equal(M, g["M"].astype(float))
    if name == "nmc_run_c1":
        assert np.array_equal(E, g["E"]) and mn == float(g["min_energy"])
    else:
        np.testing.assert_allclose(E, g["E"], rtol=1e-9)


def test_apt_preprocessor():
    g = golden("apt_preprocessor_c2")
    a = g["args"]
    seed_all(int(g["seed"]))
    beta, sigma = O.apt_preprocessor_run(g["J"], g["h"], int(a[0]), int(a[1]), int(a[2]), a[3], a[4], a[5], a[6])
    assert np.array_equal(np.array(beta, dtype=float), g["beta"])
    assert np.array_equal(np.array(sigma, dtype=float), g["sigma"])


def test_npt_run_with_nmc_replicas():
    from oracle.make_golden import NPT_KW
    g = golden("npt_run_c2")
    seed_all(int(g["seed"]))
    M, E = O.npt_run(g["J"], g["h"], g["beta_list"], 4, list(g["doNMC"]), int(g["num_sweeps_MCMC"]),
                     int(g["num_sweeps_read"]), int(g["num_swap_attempts"]), int(g["num_swapping_pairs"]), **NPT_KW)
    assert np.array_equal(M, g["M"].astype(float))
    assert np.array_equal(E, g["E"])


def test_npt_run_sk_gaussian():
    g = golden("npt_run_c3")
    seed_all(int(g["seed"]))
    M, E = O.npt_run(g["J"], g["h"], g["beta_list"], 5, [False] * 5, int(g["num_sweeps_MCMC"]),
                     int(g["num_sweeps_read"]), int(g["num_swap_attempts"]), int(g["num_swapping_pairs"]))
    assert np.array_equal(M, g["M"].astype(float))
    np.testing.assert_allclose(E, g["E"], rtol=1e-9)


def test_npt_run_sparse_input():
    g = golden("npt_run_c5")
    A, h = O.ea3d_pm_j(int(g["L"]), int(g["instance_seed"]))
    seed_all(int(g["seed"]))
    M, E = O.npt_run(A, h, g["beta_list"], 6, [False] * 6, int(g["num_sweeps_MCMC"]), int(g["num_sweeps_read"]),
                     int(g["num_swap_attempts"]), int(g["num_swapping_pairs"]))
    assert np.array_equal(M, g["M"].astype(float))
    assert np.array_equal(E, g["E"])


@pytest.mark.parametrize("tag", ["a", "b"])
def test_apt_icm_run(tag):
    g = golden("apt_icm_c4")
    nsm, nsr, nsa, npairs = (int(v) for v in g[f"{tag}_args"])
    seed_all(int(g[f"{tag}_seed"]))
    M, E = O.apt_icm_run(g["J"], g["h"], g["beta_list"], 4, nsm, nsr, nsa, npairs)
    assert np.array_equal(M, g[f"{tag}_M"].astype(float))
    assert np.array_equal(E, g[f"{tag}_E"])


def test_disagreement_clusters_element():
    g = golden("apt_icm_c4")
    labels, k = O.disagreement_clusters(O.Csr(g["J"]), g["s1"], g["s2"])
    assert k == int(g["n_clusters"])
    assert np.array_equal(labels, g["labels"])


def test_known_answer_wishart_ground_states():
    """Energy function and sign/normalisation conventions: brute force over
    2^10 states with the oracle's energy must give the shipped ground-state energy."""
    g = golden("known_answers_wishart")
    states = np.array(list(itertools.product([-1, 1], repeat=10)), dtype=np.int8)
    for J, gs in zip(g["J"], g["gs_energy"]):
        norm = np.max(np.abs(J))
        E = O.energy(O.Csr(J / norm), np.zeros(10), states)
        assert np.isclose(E.min() * norm, gs, rtol=0, atol=1e-9)


# ------------------------------------------------------------------ public element-level methods
@pytest.mark.parametrize("tag,field,warm", [("lbp", "lbp_field1", False), ("lbp2", "lbp_field2", True)])
def test_lbp_dense_call_and_byproducts(tag, field, warm):
    """LoopyBeliefPropagation's full return tuple (NMC/nmc.py:168-228): marginals, correlations, h_tilde, J_tilde,
    iteration and both message matrices."""
    g = golden("public_methods")
    J, ms = g["J"], g["lbp_m_star"].astype(float)
    n = len(ms)
    h0, u0 = (g["lbp_h_msgs"], g["lbp_u_msgs"]) if warm else (np.zeros((n, n)), J * ms.reshape(1, -1))
    marg, corr, ht, jt, it, H, U = O.lbp_dense(J, g[field], float(g["lbp_beta"]), h0, u0, float(g["lbp_tol"]),
                                               int(g["lbp_max_iter"]))
    assert it == int(g[f"{tag}_iteration"])
    assert np.array_equal(marg, g[f"{tag}_marg"])
    assert np.array_equal(H, g[f"{tag}_h_msgs"]) and np.array_equal(U, g[f"{tag}_u_msgs"])
    np.testing.assert_allclose(corr, g[f"{tag}_corr"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(ht, g[f"{tag}_h_tilde"], rtol=1e-13)
    np.testing.assert_allclose(jt, g[f"{tag}_J_tilde"], rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize("variant", ["nmc", "npt"])
def test_nmc_subroutine_with_provided_clusters(variant):
    g = golden("public_methods")
    a = g["sub_args"]
    seed_all(int(g["sub_seed"]))
    M, E, mn, cl = O.nmc_subroutine(O.Csr(g["J"]), g["h"], g["lbp_m_star"].astype(float), int(a[0]), int(a[1]),
                                    int(a[2]), int(a[3]), a[4], a[5], a[6], a[7], a[8], a[9], a[10], int(a[11]), a[12],
                                    variant, all_clusters=g["sub_clusters"])
    assert np.array_equal(M, g[f"sub_{variant}_M"])
    np.testing.assert_allclose(E, g[f"sub_{variant}_E"], rtol=1e-12)
    assert np.array_equal(cl, g[f"sub_{variant}_clusters"])
