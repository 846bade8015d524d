"""GPU parity tests (exact-replay mode): the CUDA path, called through the C ABI of libnlmc_b200.so,
against (a) the golden vectors produced by the unmodified reference and (b) the CPU oracle on seeded
inputs.  Bit-exact for spins and for integer (+-J) energies; 1e-9 relative for real-valued J
(north_star tolerance)."""
import random

import numpy as np
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu
EPS = np.finfo(float).eps


def seed_all(s):
    np.random.seed(s)
    random.seed(s)


@pytest.fixture(scope="module")
def nl():
    import nlmc_b200
    from nlmc_b200 import _lib, host, nmc_core
    return type("NL", (), dict(pkg=nlmc_b200, lib=_lib, host=host, core=nmc_core))


# ------------------------------------------------------------------ element level: K1 + K4
@pytest.mark.parametrize("tag", ["pm_fixed", "pm_anneal", "gauss_fixed", "gauss_anneal"])
def test_k1_mcmc_golden(nl, tag):
    g = golden("mcmc_element")
    J, h = g[f"{tag}_J"], g[f"{tag}_h"]
    prob = nl.host.Problem(J, h)
    reps = nl.lib.Replicas(prob.inst, 1)
    seed_all(int(g[f"{tag}_seed"]))
    m0 = np.sign(2 * np.random.rand(len(h)) - 1)
    sched = nl.host.beta_schedule(int(g[f"{tag}_sweeps"]), float(g[f"{tag}_beta"]), bool(g[f"{tag}_anneal"]), 1, 0)
    M, E = nl.host.replay_chains(prob, reps, m0[None], sched[None], np.random)
    assert np.array_equal(M[0].T, g[f"{tag}_M"])
    if tag.startswith("pm"):
        assert np.array_equal(E[0], g[f"{tag}_E"])
    else:
        np.testing.assert_allclose(E[0], g[f"{tag}_E"], rtol=1e-9)
    # K4 on the recorded states, and the final state left on the device
    E4 = prob.inst.energy_states(M[0])
    np.testing.assert_allclose(E4, g[f"{tag}_E"], rtol=1e-9)
    assert np.array_equal(reps.get_spins()[0], M[0][-1])
    np.testing.assert_allclose(reps.energy()[0], g[f"{tag}_E"][-1], rtol=1e-9)


@pytest.mark.parametrize("case", ["ea_L6", "random_graph", "sk", "field", "global_spins"])
def test_k1_vs_oracle_many_replicas(nl, case):
    """Several replicas, phase settings (scaled rows + frozen spins), zeros in the state, n > 200 KiB."""
    from oracle import oracle as O
    rs = np.random.RandomState(7)
    if case == "ea_L6":
        A, h = O.ea3d_pm_j(6, 11); J = A
    elif case == "random_graph":
        J, h = O.random_pm_graph(200, 0.06, 12)
    elif case == "sk":
        J, h = O.sk_gaussian(96, 13)
    elif case == "field":
        J, h = O.random_pm_graph(80, 0.1, 14); h = rs.randn(80) * 0.3
    else:  # spins live in global memory when n > 200 KiB
        A, h = O.ea3d_pm_j(60, 15); J = A
    csr = O.Csr(J)
    n = csr.n
    R, S = (3, 2) if case == "global_spins" else (5, 4)
    prob = nl.host.Problem(J, h)
    reps = nl.lib.Replicas(prob.inst, R)
    m0 = rs.choice([-1, 1], size=(R, n)).astype(np.int8)
    m0[0, :3] = 0  # np.sign can produce 0; the kernel must carry it
    betas = np.linspace(0.3, 2.0, R)
    sched = np.repeat(betas[:, None], S, axis=1)
    perm = np.stack([np.stack([rs.permutation(n) for _ in range(S)]) for _ in range(R)]).astype(np.int32)
    u = rs.rand(R, S, n)
    h_eff, scaled = [None] * R, [None] * R
    if case in ("random_graph", "sk", "field"):
        for r in (1, 3):
            in_cl = rs.rand(n) < 0.3
            he = np.asarray(h, dtype=float).copy()
            he[in_cl] /= 20
            he[~in_cl] = m0[r][~in_cl] * 10000.0
            h_eff[r], scaled[r] = he, in_cl
            reps.set_phase(r, he, in_cl.astype(np.uint8), 20)
    reps.set_spins(m0)
    M, E = reps.sweep_replay(perm, u, sched, prob.tanh_lut(sched), prob.lut_half)
    rows = csr.row_of
    for r in range(R):
        c = csr if scaled[r] is None else csr.with_values(np.where(scaled[r][rows], csr.val / 20, csr.val))
        Mo, _ = O.mcmc(c, h if h_eff[r] is None else h_eff[r], m0[r], sched[r], perm=perm[r], u=u[r])
        assert np.array_equal(M[r], Mo), f"replica {r}"
        Eo = O.energy(csr, h, Mo)
        if prob.is_integer and not np.any(np.asarray(h) % 1):
            assert np.array_equal(E[r], Eo)
        else:
            np.testing.assert_allclose(E[r], Eo, rtol=1e-9, atol=1e-12)


def test_k1_record_from_and_empty(nl):
    from oracle import oracle as O
    J, h = O.random_pm_graph(30, 0.2, 5)
    prob = nl.host.Problem(J, h)
    reps = nl.lib.Replicas(prob.inst, 2)
    rs = np.random.RandomState(1)
    m0 = rs.choice([-1, 1], size=(2, 30)).astype(np.int8)
    reps.set_spins(m0)
    perm = np.stack([np.stack([rs.permutation(30) for _ in range(5)]) for _ in range(2)]).astype(np.int32)
    u = rs.rand(2, 5, 30)
    sched = np.full((2, 5), 0.9)
    M_all, E_all = reps.sweep_replay(perm, u, sched, None, 0, record_from=0)
    reps.set_spins(m0)
    M_tail, _ = reps.sweep_replay(perm, u, sched, None, 0, record_from=3)
    assert np.array_equal(M_tail, M_all[:, 3:])
    # zero sweeps: nothing happens, nothing is returned
    M0, E0 = reps.sweep_replay(np.zeros((2, 0, 30), np.int32), np.zeros((2, 0, 30)), np.zeros((2, 0)))
    assert M0.shape == (2, 0, 30) and E0.shape == (2, 0)
    assert np.array_equal(reps.get_spins(), M_all[:, -1])


# ------------------------------------------------------------------ K5: LBP
def test_k5_lbp_golden(nl):
    """Every lambda step of the reference's LBP_convexified (tolerance = machine epsilon): the same iteration
    count, bit-identical marginals, the same divergence point and the same backbone -- unconditionally; the
    device tanh/arctanh are bit-equal to numpy's (tests/test_gpu_npmath.py)."""
    g = golden("lbp")
    J, h, ms = g["J"], g["h"], g["m_star"].astype(float)
    prob = nl.host.Problem(J, h)
    lbp = nl.lib.Lbp(prob.inst)
    assert np.array_equal(lbp.epsilon(), np.abs(h) + np.sum(np.abs(J), axis=1))
    lam0, lam_end, fac, tol, max_it, thr_i, thr_c = g["params"]
    lbp.reset(ms)
    assert len(g["lambdas"]) >= 10
    for lam, marg_ref, it_ref in zip(g["lambdas"], g["marginals"], g["iters"]):
        marg, it = lbp.step(lam, float(g["beta"]), tol, int(max_it))
        assert it == it_ref, (lam, it, it_ref)
        if it_ref != int(max_it) - 1:
            assert np.array_equal(marg, marg_ref), lam
    trace = []
    cl = nl.core.lbp_convexified(prob, lbp, ms, lam0, lam_end, fac, tol, int(max_it), thr_i, thr_c, float(g["beta"]),
                                 trace=trace)
    assert [t[1] for t in trace] == list(g["iters"])
    assert np.array_equal(np.concatenate(cl), g["clusters"])


def test_k5_lbp_vs_oracle_gaussian(nl):
    from oracle import oracle as O
    J, h = O.sk_gaussian(50, 21)
    rs = np.random.RandomState(3)
    h = rs.randn(50) * 0.1
    ms = rs.choice([-1.0, 1.0], size=50)
    csr = O.Csr(J)
    prob = nl.host.Problem(J, h)
    lbp = nl.lib.Lbp(prob.inst)
    eps_o = np.abs(h) + O._pairwise_rowsum_abs(csr)
    assert np.array_equal(lbp.epsilon(), eps_o)  # pairwise row sums reproduced bit for bit
    lbp.reset(ms)
    u = np.ascontiguousarray(csr.val * ms[csr.ci]); hm = np.zeros_like(u); tot = np.zeros(50)
    lam = 3.0
    for _ in range(6):
        marg_o, it_o = O.lbp(csr, np.ascontiguousarray(h + lam * ms * eps_o), 1.5, u, hm, tot, 1e-10, 200)
        marg, it = lbp.step(lam, 1.5, 1e-10, 200)
        assert it == it_o
        assert np.array_equal(marg, marg_o)   # the oracle uses numpy's own tanh/arctanh: bit-identical marginals
        lam *= 0.9


# ------------------------------------------------------------------ K7: ICM clusters
def test_k7_clusters_golden_and_oracle(nl):
    from oracle import oracle as O
    g = golden("apt_icm_c4")
    prob = nl.host.Problem(g["J"], g["h"])
    labels, counts = nl.lib.icm_clusters(prob.inst, g["s1"], g["s2"])
    assert counts[0] == int(g["n_clusters"]) and np.array_equal(labels[0], g["labels"])
    # larger lattice, many pairs in one launch, including identical and opposite states
    A, h = O.ea3d_pm_j(12, 3)
    csr = O.Csr(A)
    prob = nl.host.Problem(A, h)
    rs = np.random.RandomState(5)
    s1 = rs.choice([-1, 1], size=(9, csr.n)).astype(np.int8)
    s2 = np.where(rs.rand(9, csr.n) < np.linspace(0.05, 0.9, 9)[:, None], -s1, s1).astype(np.int8)
    s2[0] = s1[0]          # no disagreement -> no clusters
    s2[8] = -s1[8]         # full disagreement -> one cluster
    labels, counts = nl.lib.icm_clusters(prob.inst, s1, s2)
    for p in range(9):
        lo, ko = O.disagreement_clusters(csr, s1[p], s2[p])
        assert counts[p] == ko and np.array_equal(labels[p], lo)
    assert counts[0] == 0 and counts[8] == 1


# ------------------------------------------------------------------ run() level vs golden
def test_npt_run_golden_sparse_input(nl, tmp_cwd):
    from oracle import oracle as O
    g = golden("npt_run_c5")
    A, h = O.ea3d_pm_j(int(g["L"]), int(g["instance_seed"]))
    seed_all(int(g["seed"]))
    M, E = nl.pkg.NPT(A, h).run(g["beta_list"], 6, [False] * 6, num_sweeps_MCMC=int(g["num_sweeps_MCMC"]),
                                num_sweeps_read=int(g["num_sweeps_read"]),
                                num_swap_attempts=int(g["num_swap_attempts"]),
                                num_swapping_pairs=int(g["num_swapping_pairs"]), num_cores=1)
    assert M.dtype == np.float64 and np.array_equal(M, g["M"].astype(float))
    assert np.array_equal(E, g["E"])


def test_npt_run_golden_sk(nl, tmp_cwd):
    g = golden("npt_run_c3")
    seed_all(int(g["seed"]))
    M, E = nl.pkg.NPT(g["J"], g["h"]).run(g["beta_list"], 5, [False] * 5, num_sweeps_MCMC=int(g["num_sweeps_MCMC"]),
                                          num_sweeps_read=int(g["num_sweeps_read"]),
                                          num_swap_attempts=int(g["num_swap_attempts"]),
                                          num_swapping_pairs=int(g["num_swapping_pairs"]), num_cores=1)
    assert np.array_equal(M, g["M"].astype(float))
    np.testing.assert_allclose(E, g["E"], rtol=1e-9)


def test_npt_run_golden_with_nmc_replicas(nl, tmp_cwd):
    """C2-shaped NPT.run with doNMC replicas, free-running (K5 finds the backbones): bit-exact."""
    from oracle.make_golden import NPT_KW
    g = golden("npt_run_c2")
    seed_all(int(g["seed"]))
    M, E = nl.pkg.NPT(g["J"], g["h"]).run(g["beta_list"], 4, list(g["doNMC"]),
                                          num_sweeps_MCMC=int(g["num_sweeps_MCMC"]),
                                          num_sweeps_read=int(g["num_sweeps_read"]),
                                          num_swap_attempts=int(g["num_swap_attempts"]),
                                          num_swapping_pairs=int(g["num_swapping_pairs"]), num_cores=1, **NPT_KW)
    assert np.array_equal(M, g["M"].astype(float))
    assert np.array_equal(E, g["E"])


def test_apt_preprocessor_golden(nl, tmp_cwd):
    import os
    g = golden("apt_preprocessor_c2")
    a = g["args"]
    seed_all(int(g["seed"]))
    beta, sigma = nl.pkg.APT_preprocessor(g["J"].copy(), g["h"].copy()).run(
        num_sweeps_MCMC=int(a[0]), num_sweeps_read=int(a[1]), num_rng=int(a[2]), beta_start=a[3], alpha=a[4],
        sigma_E_val=a[5], beta_max=a[6], use_hash_table=0, num_cores=1)
    assert isinstance(beta, list) and isinstance(sigma, list)
    assert np.array_equal(np.array(beta, dtype=float), g["beta"])
    assert np.array_equal(np.array(sigma, dtype=float), g["sigma"])
    assert os.path.exists("beta_list_python.npy") and os.path.exists("Results/data/Energy_iter_1.npy")


@pytest.mark.parametrize("tag", ["a", "b"])
def test_apt_icm_golden(nl, tmp_cwd, tag):
    g = golden("apt_icm_c4")
    nsm, nsr, nsa, npairs = (int(v) for v in g[f"{tag}_args"])
    seed_all(int(g[f"{tag}_seed"]))
    obj = nl.pkg.APT_ICM(g["J"].copy(), g["h"].copy())
    M, E = obj.run(g["beta_list"], 4, num_sweeps_MCMC=nsm, num_sweeps_read=nsr, num_swap_attempts=nsa,
                   num_swapping_pairs=npairs)
    assert obj.num_sweeps_MCMC == nsm
    assert np.array_equal(M, g[f"{tag}_M"].astype(float))
    assert np.array_equal(E, g[f"{tag}_E"])


@pytest.mark.parametrize("name", ["nmc_run_c1", "nmc_run_gauss"])
def test_nmc_run_golden(nl, tmp_cwd, name):
    """Whole NMC.run (C1-shaped and the reference's unit-test shape), free-running: anneal, ten LBP backbone
    searches and every NMC phase reproduce the reference's states bit for bit."""
    g = golden(name)
    a = g["args"]
    args = (int(a[0]), int(a[1]), int(a[2]), int(a[3]), int(a[4]), a[5], a[6], a[7], a[8], a[9], a[10], a[11],
            int(a[12]), a[13])
    seed_all(int(g["seed"]))
    M, E, mn = nl.pkg.NMC(g["J"], g["h"]).run(*args)
    assert isinstance(M, np.ndarray) and M.shape == g["M"].shape
    assert isinstance(mn, (float, np.float64))
    assert np.array_equal(M, g["M"].astype(float))
    np.testing.assert_allclose(E, g["E"], rtol=1e-9)
    np.testing.assert_allclose(mn, float(g["min_energy"]), rtol=1e-9)


def test_k1_int_kernel_equals_general_kernel(nl, monkeypatch):
    """The shared-memory incremental-field kernel (integer J) and the general kernel must agree bit for bit,
    including NMC phase settings (rescaled rows, frozen spins), a real-valued h and zeros in the state."""
    from oracle import oracle as O
    rs = np.random.RandomState(11)
    J, h = O.random_pm_graph(300, 0.08, 31)
    J[5, 9] = J[9, 5] = 3.0  # an integer coupling other than +-1
    h = np.where(rs.rand(300) < 0.2, rs.randn(300), 0.0)
    prob = nl.host.Problem(J, h)
    R, S, n = 6, 5, 300
    m0 = rs.choice([-1, 1], size=(R, n)).astype(np.int8)
    m0[2, :4] = 0
    sched = np.repeat(np.linspace(0.2, 3.0, R)[:, None], S, axis=1)
    perm = np.stack([np.stack([rs.permutation(n) for _ in range(S)]) for _ in range(R)]).astype(np.int32)
    u = rs.rand(R, S, n)
    outs = []
    for force_general in (False, True):
        if force_general:
            monkeypatch.setenv("NLMC_REPLAY_GENERAL", "1")
        reps = nl.lib.Replicas(prob.inst, R, m0)
        for r in (1, 4):
            in_cl = np.random.RandomState(r).rand(n) < 0.3
            he = h.copy()
            he[in_cl] /= 20
            he[~in_cl] = m0[r][~in_cl] * 10000.0
            reps.set_phase(r, he, in_cl.astype(np.uint8), 20)
        M, E = reps.sweep_replay(perm, u, sched, prob.tanh_lut(sched), prob.lut_half)
        outs.append((M, E, reps.get_spins()))
    monkeypatch.delenv("NLMC_REPLAY_GENERAL")
    assert np.array_equal(outs[0][0], outs[1][0])
    np.testing.assert_allclose(outs[0][1], outs[1][1], rtol=1e-12, atol=1e-9)
    assert np.array_equal(outs[0][2], outs[1][2])
    # and both equal the oracle
    csr = O.Csr(J)
    Mo, _ = O.mcmc(csr, h, m0[0], sched[0], perm=perm[0], u=u[0])
    assert np.array_equal(outs[0][0][0], Mo)


@pytest.mark.parametrize("n,p,mode", [(300, 0.1, "warp"), (1100, 0.005, "thread"), (2500, 0.02, "warp"), (129, 0.5, "warp")])
def test_k5_gather_reproduces_numpy_summation_order_beyond_one_block(nl, n, p, mode):
    """numpy's pairwise sum recurses for more than 128 elements; the per-row summation programs K5 replays (resolved on
    the host from the sparsity pattern) must reproduce that association exactly.  After ONE iteration from identical
    inputs the column totals and the h messages are pure additions -- compared bit for bit with the oracle, which uses
    numpy's own order -- on sizes that need 2-5 levels of the recursion, in both gather modes."""
    from oracle import oracle as O
    rs = np.random.RandomState(n)
    iu = np.triu_indices(n, 1)
    keep = rs.rand(len(iu[0])) < p
    J = np.zeros((n, n))
    J[iu[0][keep], iu[1][keep]] = rs.randn(int(keep.sum()))
    J += J.T
    h = 0.1 * rs.randn(n)
    ms = rs.choice([-1.0, 1.0], size=n)
    csr = O.Csr(J)
    prob = nl.host.Problem(J, h)
    assert (len(prob.val) >= 12 * n) == (mode == "warp")
    lbp = nl.lib.Lbp(prob.inst)
    eps_o = np.abs(h) + O._pairwise_rowsum_abs(csr)
    assert np.array_equal(lbp.epsilon(), eps_o)
    assert np.array_equal(eps_o, np.abs(h) + np.sum(np.abs(J), axis=1))        # the oracle itself against numpy
    lbp.reset(ms)
    lbp.step(0.8, 1.3, -1.0, 1)                                                  # tolerance < 0: exactly one iteration
    hm, _, tot = lbp.get_messages()
    u = np.ascontiguousarray(csr.val * ms[csr.ci]); hm_o = np.zeros_like(u); tot_o = np.zeros(n)
    O.lbp(csr, np.ascontiguousarray(h + 0.8 * ms * eps_o), 1.3, u, hm_o, tot_o, -1.0, 1)
    assert np.array_equal(tot, tot_o) and np.array_equal(hm, hm_o)
    # and the dense numpy expression itself for the totals: h_lambda + np.sum(u_msgs[:, i]) over the strided column
    U0 = J * ms.reshape(1, -1)
    ref = np.array([(h + 0.8 * ms * eps_o)[i] + np.sum(U0[:, i]) for i in range(n)])
    assert np.array_equal(tot, ref)


def test_k7_clusters_at_config_c4_size(nl):
    """Houdayer disagreement clusters at the size of config C4 (3D EA L = 32, 32,768 spins) and on a random graph of
    degree ~48: labels and cluster order identical to the oracle's breadth-first search."""
    from nlmc_b200 import instances
    from oracle import oracle as O
    for J, h in (instances.ea3d_pm_j(32, 4), instances.random_pm_graph(800, 0.06, 1)):
        csr = O.Csr(J)
        prob = nl.host.Problem(J, h)
        rs = np.random.RandomState(11)
        s1 = rs.choice([-1, 1], size=(4, csr.n)).astype(np.int8)
        s2 = np.where(rs.rand(4, csr.n) < np.array([0.02, 0.2, 0.45, 0.7])[:, None], -s1, s1).astype(np.int8)
        labels, counts = nl.lib.icm_clusters(prob.inst, s1, s2)
        for p in range(4):
            lo, ko = O.disagreement_clusters(csr, s1[p], s2[p])
            assert counts[p] == ko and np.array_equal(labels[p], lo), (csr.n, p)


@pytest.mark.parametrize("config", ["C1", "C2", "C4"])
def test_k1_replay_at_config_sizes(nl, config):
    """K1 against the oracle at the native sizes of the configs (C1: N = 800 random +-1 graph of degree ~48; C2: 3D EA
    L = 16; C4: 3D EA L = 32), a plain replica and one in an NMC phase (rows at beta/temp_x, the rest frozen)."""
    from nlmc_b200 import instances
    from oracle import oracle as O
    J, h = {"C1": lambda: instances.random_pm_graph(800, 0.06, 1), "C2": lambda: instances.ea3d_pm_j(16, 2),
            "C4": lambda: instances.ea3d_pm_j(32, 4)}[config]()
    csr = O.Csr(J)
    n = csr.n
    prob = nl.host.Problem(J, h)
    rs = np.random.RandomState(17)
    R, S = 2, 3
    reps = nl.lib.Replicas(prob.inst, R)
    m0 = rs.choice([-1, 1], size=(R, n)).astype(np.int8)
    sched = np.array([[0.7] * S, [2.5] * S])
    perm = np.stack([np.stack([rs.permutation(n) for _ in range(S)]) for _ in range(R)]).astype(np.int32)
    u = rs.rand(R, S, n)
    in_cl = rs.rand(n) < 0.6
    he = np.zeros(n)
    he[~in_cl] = m0[1][~in_cl] * 10000.0
    reps.set_phase(1, he, in_cl.astype(np.uint8), 20)
    reps.set_spins(m0)
    M, E = reps.sweep_replay(perm, u, sched, prob.tanh_lut(sched), prob.lut_half)
    Mo0, _ = O.mcmc(csr, h, m0[0], sched[0], perm=perm[0], u=u[0])
    c1 = csr.with_values(np.where(in_cl[csr.row_of], csr.val / 20, csr.val))
    Mo1, _ = O.mcmc(c1, he, m0[1], sched[1], perm=perm[1], u=u[1])
    assert np.array_equal(M[0], Mo0) and np.array_equal(M[1], Mo1)
    assert np.array_equal(E[0], O.energy(csr, h, Mo0)) and np.array_equal(E[1], O.energy(csr, h, Mo1))


@pytest.mark.parametrize("config", ["C3", "C5"])
def test_k1_replay_at_config_sizes_c3_c5(nl, config):
    """K1 against the oracle at the native sizes of C3 (SK N = 2000, dense Gaussian J: storage-order fp64 row sums over
    1999 entries, energies to 1e-9) and C5 (3D EA L = 64, 262,144 spins: global-memory variant), a plain replica and
    one in an NMC phase (rows at beta/temp_x, the rest frozen)."""
    from nlmc_b200 import instances
    from oracle import oracle as O
    J, h = instances.sk_gaussian(2000, 3) if config == "C3" else instances.ea3d_pm_j(64, 5)
    csr = O.Csr(J)
    n = csr.n
    prob = nl.host.Problem(J, h)
    rs = np.random.RandomState(23)
    R, S = 2, (2 if config == "C3" else 1)
    reps = nl.lib.Replicas(prob.inst, R)
    m0 = rs.choice([-1, 1], size=(R, n)).astype(np.int8)
    sched = np.array([[0.9] * S, [2.5] * S])
    perm = np.stack([np.stack([rs.permutation(n) for _ in range(S)]) for _ in range(R)]).astype(np.int32)
    u = rs.rand(R, S, n)
    in_cl = rs.rand(n) < 0.5
    he = np.zeros(n)
    he[~in_cl] = m0[1][~in_cl] * 10000.0
    reps.set_phase(1, he, in_cl.astype(np.uint8), 20)
    reps.set_spins(m0)
    lut = prob.tanh_lut(sched) if prob.lut_half else None
    M, E = reps.sweep_replay(perm, u, sched, lut, prob.lut_half)
    Mo0, _ = O.mcmc(csr, h, m0[0], sched[0], perm=perm[0], u=u[0])
    c1 = csr.with_values(np.where(in_cl[csr.row_of], csr.val / 20, csr.val))
    Mo1, _ = O.mcmc(c1, he, m0[1], sched[1], perm=perm[1], u=u[1])
    assert np.array_equal(M[0], Mo0) and np.array_equal(M[1], Mo1)
    if config == "C5":
        assert np.array_equal(E[0], O.energy(csr, h, Mo0)) and np.array_equal(E[1], O.energy(csr, h, Mo1))
    else:
        np.testing.assert_allclose(E[0], O.energy(csr, h, Mo0), rtol=1e-9)
        np.testing.assert_allclose(E[1], O.energy(csr, h, Mo1), rtol=1e-9)


def test_nmc_run_at_config_c1_size_vs_oracle(nl, tmp_cwd):
    """Whole NMC.run on the C1 instance (N = 800, +-1 graph of degree ~48, README parameters, sweeps cut to 20 and
    3 cycles) against the oracle's nmc_run on the same seeds: anneal, three free-running LBP backbone searches at
    tolerance = machine epsilon, nine phases -- states bit for bit, energies exact (integer J)."""
    from nlmc_b200 import instances
    from oracle import oracle as O
    J, h = instances.random_pm_graph(800, 0.06, 1)
    args = (20, 20, 3, 1, 1, 20, 3, 3, 0.01, 0.9, 0.9999999, 0.999999, 100, EPS)
    seed_all(3)
    Mo, Eo, mno = O.nmc_run(J, h, *args)
    seed_all(3)
    M, E, mn = nl.pkg.NMC(J.toarray() if hasattr(J, "toarray") else J, h).run(*args)
    assert M.shape == (800, 180)
    assert np.array_equal(M, Mo)
    assert np.array_equal(np.asarray(E), Eo) and mn == mno


def test_npt_run_at_config_c2_size_vs_oracle(nl, tmp_cwd):
    """Whole NPT.run on the C2 lattice (3D EA L = 16, 4096 spins), 6 replicas with doNMC on the 2 coldest, against
    the oracle's npt_run on the same seeds (free-running LBP, swaps, forked worker stream): bit for bit."""
    from nlmc_b200 import instances
    from oracle import oracle as O
    J, h = instances.ea3d_pm_j(16, 2)
    bl = np.linspace(0.3, 2.5, 6)
    kw = dict(num_sweeps_MCMC=24, num_sweeps_read=24, num_swap_attempts=2, num_swapping_pairs=2, num_cycles=2,
              global_beta=3.0, lambda_start=3.0, lambda_end=0.01, threshold_initial=0.9999999,
              threshold_cutoff=0.999999)
    doNMC = [False] * 4 + [True] * 2
    seed_all(4)
    Mo, Eo = O.npt_run(J, h, bl, 6, doNMC, **kw)
    seed_all(4)
    M, E = nl.pkg.NPT(J, h).run(bl, 6, doNMC, num_cores=1, **kw)
    assert M.shape == (6 * 4096, 12)
    assert np.array_equal(M, Mo) and np.array_equal(E, Eo)
