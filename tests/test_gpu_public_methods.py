"""GPU parity of the classes' element-level public methods (MCMC, LoopyBeliefPropagation, LBP_convexified,
NMC_subroutine, MCMC_task, NMC_task, replica_energy, find_clusters, atanh_saturated) against golden vectors the
unmodified reference produced (oracle/make_golden.py: case_public_methods).  Sweeps and energies are bit-exact;
belief propagation is compared at 1e-9 (CUDA's tanh/atanh differ from numpy's in the last place; tolerance chosen
two orders above the LBP stopping tolerance of the fixtures)."""
import random

import numpy as np
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu
EPS = np.finfo(float).eps


def seed_all(s):
    np.random.seed(s)
    random.seed(s)


def assert_close(got, ref, beta, key, tol=1e-8):
    """h_tilde / J_tilde are atanh(x)/beta of a marginal / correlation x; near saturation (|x| -> 1) atanh amplifies
    a last-place difference of x by 1/(1-x^2) ~ 1e7, so they are compared as tanh(beta * value), i.e. on x itself."""
    if key in ("h_tilde", "J_tilde"):
        got, ref = np.tanh(beta * got), np.tanh(beta * ref)
    np.testing.assert_allclose(got, ref, rtol=tol, atol=tol, err_msg=key)


@pytest.fixture(scope="module")
def g():
    return golden("public_methods")


@pytest.fixture(scope="module")
def pkg():
    import nlmc_b200
    return nlmc_b200


@pytest.mark.parametrize("tag", ["pm_fixed", "pm_anneal", "gauss_fixed", "gauss_anneal"])
def test_mcmc_method_matches_reference(pkg, tag):
    """NMC.MCMC / NPT.MCMC called the way the reference's own code calls them (NMC/nmc.py:490, NPT/npt.py:124)."""
    e = golden("mcmc_element")
    J, h = e[f"{tag}_J"], e[f"{tag}_h"]
    for cls in (pkg.NMC, pkg.NPT):
        obj = cls(J, h)
        seed_all(int(e[f"{tag}_seed"]))
        m0 = np.sign(2 * np.random.rand(len(h)) - 1)
        M = obj.MCMC(int(e[f"{tag}_sweeps"]), m0.copy(), float(e[f"{tag}_beta"]), J, h, anneal=bool(e[f"{tag}_anneal"]))
        assert M.dtype == np.float64 and M.shape == e[f"{tag}_M"].shape
        assert np.array_equal(M, e[f"{tag}_M"])


def test_mcmc_method_edge_cases(pkg, g):
    J, h = g["J"], g["h"]
    obj = pkg.NMC(J, h)
    m0 = g["lbp_m_star"].astype(float)
    assert obj.MCMC(0, m0, 1.0, J, h).shape == (len(h), 0)
    with pytest.raises(ValueError):
        obj.MCMC(-1, m0, 1.0, J, h)
    with pytest.raises(ValueError, match="LRUCache"):
        obj.MCMC(2, m0, 1.0, J, h, hash_table={}, use_hash_table=True)
    # (N,1)-shaped start and h, sparse J: same stream, same result
    import scipy.sparse as sp
    seed_all(5)
    A = obj.MCMC(3, m0.reshape(-1, 1), 1.3, sp.csr_matrix(J), h.reshape(-1, 1))
    seed_all(5)
    B = obj.MCMC(3, m0, 1.3, J, h)
    assert np.array_equal(A, B)
    # the caller's arrays are not modified
    assert np.array_equal(m0, g["lbp_m_star"])


@pytest.mark.parametrize("tag,field,warm", [("lbp", "lbp_field1", False), ("lbp2", "lbp_field2", True)])
def test_loopy_belief_propagation_full_tuple(pkg, g, tag, field, warm):
    J, ms = g["J"], g["lbp_m_star"].astype(float)
    n = len(ms)
    obj = pkg.NMC(J, g["h"])
    h0, u0 = (g["lbp_h_msgs"], g["lbp_u_msgs"]) if warm else (np.zeros((n, n)), J * ms.reshape(1, -1))
    h0c, u0c = h0.copy(), u0.copy()
    marg, corr, ht, jt, it, H, U = obj.LoopyBeliefPropagation(J, g[field], float(g["lbp_beta"]), h0, u0,
                                                              float(g["lbp_tol"]), int(g["lbp_max_iter"]))
    assert np.array_equal(h0, h0c) and np.array_equal(u0, u0c)
    assert abs(it - int(g[f"{tag}_iteration"])) <= 1   # the stopping rule compares against 1e-10
    for got, key in ((marg, "marg"), (corr, "corr"), (ht, "h_tilde"), (jt, "J_tilde"), (H, "h_msgs"), (U, "u_msgs")):
        assert got.shape == g[f"{tag}_{key}"].shape
        assert_close(got, g[f"{tag}_{key}"], float(g["lbp_beta"]), key)


def test_loopy_belief_propagation_arbitrary_dense_messages(pkg, g):
    """Messages that are non-zero off the entries of J (never produced by the reference itself, but legal input)."""
    obj = pkg.NPT(g["J"], g["h"])
    out = obj.LoopyBeliefPropagation(g["J"], g["h"].copy(), 1.5, g["lbp3_h0"].copy(), g["lbp3_u0"].copy(), 1e-10, 3)
    assert out[4] == int(g["lbp3_iteration"])
    for got, key in zip(out, ("marg", "corr", "h_tilde", "J_tilde", None, "h_msgs", "u_msgs")):
        if key:
            assert_close(got, g[f"lbp3_{key}"], 1.5, key, tol=1e-9)


def test_lbp_convexified_dictionaries(pkg, g):
    J, h, ms = g["J"], g["h"], g["lbp_m_star"].astype(float)
    a = g["conv_args"]
    obj = pkg.NMC(J, h)
    epsv = np.abs(h) + np.sum(np.abs(J), axis=1)
    cl, marg, mean, ht, jt = obj.LBP_convexified(a[0], a[1], a[2], ms.copy(), epsv, a[3], int(a[4]), a[5], a[6], a[7])
    lams = list(marg.keys())
    np.testing.assert_array_equal(np.array(lams), g["conv_lambdas"])
    assert list(mean.keys()) == lams and list(ht.keys()) == lams and list(jt.keys()) == lams
    np.testing.assert_allclose(np.array([marg[k] for k in lams]), g["conv_marginals"], rtol=1e-7, atol=1e-7)
    np.testing.assert_allclose(np.array([mean[k] for k in lams]), g["conv_means"], rtol=1e-7, atol=1e-7)
    # by-products of every converged lambda (the last step hit max_iterations in the reference: not a fixed point)
    assert_close(np.array([ht[k] for k in lams[:-1]]), g["conv_h_tilde"][:-1], a[7], "h_tilde", tol=1e-7)
    assert jt[lams[-1]].shape == g["conv_J_tilde_last"].shape
    assert np.array_equal(np.concatenate(cl), g["conv_clusters_flat"])
    assert [len(c) for c in cl] == list(g["conv_cluster_sizes"])
    # find_clusters as a public method gives the same grouping from the final marginal
    cl2 = obj.find_clusters(marg[lams[-1]], a[5], a[6], 0.01)
    assert [list(c) for c in cl2] == [list(c) for c in cl]


def test_atanh_saturated(pkg, g):
    obj = pkg.NMC(g["J"], g["h"])
    x = np.array([-2.0, -1.0, -0.5, 0.0, 0.3, 1.0, 7.0])
    got = obj.atanh_saturated(x)
    assert np.all(np.isfinite(got)) and got[0] == got[1] and got[-1] == got[-2]
    np.testing.assert_allclose(got[2:5], np.arctanh(x[2:5]))


@pytest.mark.parametrize("variant", ["nmc", "npt"])
def test_nmc_subroutine_provided_clusters_exact(pkg, g, variant):
    a = g["sub_args"]
    obj = (pkg.NMC if variant == "nmc" else pkg.NPT)(g["J"], g["h"])
    seed_all(int(g["sub_seed"]))
    M, E, mn, cl = obj.NMC_subroutine(g["lbp_m_star"].astype(float), int(a[0]), int(a[1]), int(a[2]), int(a[3]), a[4],
                                      a[5], a[6], a[7], a[8], a[9], a[10], int(a[11]), a[12],
                                      all_clusters=g["sub_clusters"].copy())
    assert np.array_equal(M, g[f"sub_{variant}_M"])
    np.testing.assert_allclose(E, g[f"sub_{variant}_E"], rtol=1e-9)
    assert mn == pytest.approx(float(g[f"sub_{variant}_min"]), rel=1e-9)
    assert np.array_equal(cl, g[f"sub_{variant}_clusters"])


def test_npt_tasks_and_replica_energy(pkg, g):
    from nlmc_b200 import nmc_core
    obj = pkg.NPT(g["J"], g["h"])
    ms = g["lbp_m_star"].astype(float)
    seed_all(int(g["task_seed"]))
    Mt = obj.MCMC_task(2, 6, ms.copy(), g["task_betas"])
    assert np.array_equal(Mt, g["task_M"])
    mn, EE1 = obj.replica_energy(Mt, 4)
    np.testing.assert_allclose(EE1, g["rep_EE1"], rtol=1e-9)
    assert mn == pytest.approx(float(g["rep_min"]), rel=1e-9)
    # NMC_task with the backbone the reference found (LBP at tolerance 1e-9 gives the same one; checked below)
    a = g["nmctask_args"]
    args = (int(a[0]), int(a[1]), int(a[2]), int(a[3]), a[4], a[5], a[6], a[7], a[8], a[9], a[10], int(a[11]), a[12])
    seed_all(int(g["nmctask_seed"]))
    Mn = obj.NMC_task(ms.copy(), *args)
    assert np.array_equal(Mn, g["nmctask_M"])


def test_apt_classes_mcmc_and_task(pkg, g):
    J, h, ms = g["J"], g["h"], g["lbp_m_star"].astype(float)
    prep, icm = pkg.APT_preprocessor(J, h), pkg.APT_ICM(J, h)
    seed_all(int(g["apt_seed"]))
    Mp = prep.MCMC(5, ms.copy(), 1.2)
    En, mlast = prep.MCMC_task(ms.copy(), 0.8, 7, 3)
    Mi = icm.MCMC(4, ms.copy(), 0.6)
    mn_i, EE_i = icm.replica_energy(Mi, 4)
    assert np.array_equal(Mp, g["prep_M"]) and np.array_equal(Mi, g["icm_M"])
    np.testing.assert_allclose(En, g["prep_task_E"], rtol=1e-9)
    assert mlast.shape == g["prep_task_m"].shape and np.array_equal(mlast, g["prep_task_m"])
    np.testing.assert_allclose(EE_i, g["icm_rep_EE1"], rtol=1e-9)
    assert mn_i == pytest.approx(float(g["icm_rep_min"]), rel=1e-9)


def test_production_mode_methods_run(pkg, g):
    """mode='production': same methods, Philox streams; results are valid spin matrices with energies that drop."""
    J, h, ms = g["J"], g["h"], g["lbp_m_star"].astype(float)
    obj = pkg.NMC(J, h, mode="production")
    seed_all(1)
    M = obj.MCMC(40, ms.copy(), 3.0, J, h, anneal=True)
    assert M.shape == (len(h), 40) and set(np.unique(M)) <= {-1.0, 1.0}
    E = [-(M[:, i] @ J @ M[:, i] / 2 + M[:, i] @ h) for i in (0, 39)]
    assert E[1] < E[0]
    Mo, Eo, mn, cl = obj.NMC_subroutine(M[:, -1], 2, 6, 1, 2, 2.0, 10.0, 2.0, 0.05, 0.8, 0.99, 0.9, 300, 1e-9,
                                        all_clusters=g["sub_clusters"])
    assert Mo.shape == (len(h), 18) and len(Eo) == 18 and mn == Eo.min()
    np.testing.assert_allclose(Eo, [-(Mo[:, i] @ J @ Mo[:, i] / 2 + Mo[:, i] @ h) for i in range(18)], rtol=1e-9)


def test_lbp_on_csr_with_unsorted_rows(g):
    """K5 needs rows sorted by column (numpy sums the reference's dense rows and columns in index order); the C ABI
    says so, and the host hands it a sorted twin of an instance whose CSR is stored in another order -- the replay
    sweeps keep the caller's storage order (scipy's J.dot order), the backbone search is independent of it."""
    import scipy.sparse as sp
    from nlmc_b200 import _lib, host, nmc_core
    J, h, ms = g["J"], g["h"], g["lbp_m_star"].astype(float)
    A = sp.csr_matrix(J)
    rs = np.random.RandomState(0)
    indices, data = A.indices.copy(), A.data.copy()
    for i in range(A.shape[0]):
        b, e = A.indptr[i], A.indptr[i + 1]
        p = rs.permutation(e - b)
        indices[b:e], data[b:e] = indices[b:e][p], data[b:e][p]
    B = sp.csr_matrix((data, indices, A.indptr.copy()), shape=A.shape)
    pa, pb = host.Problem(A, h), host.Problem(B, h)
    assert np.array_equal(pb.ci, indices) and pa.lbp_instance() is pa.inst and pb.lbp_instance() is not pb.inst
    with pytest.raises(_lib.NlmcError, match="not sorted"):
        _lib.Lbp(pb.inst)
    a = g["conv_args"]
    outs = []
    for prob in (pa, pb):
        lbp = _lib.Lbp(prob.lbp_instance())
        trace = []
        cl = nmc_core.lbp_convexified(prob, lbp, ms, a[0], a[1], a[2], a[3], int(a[4]), a[5], a[6], a[7], trace=trace)
        outs.append((cl, trace))
        lbp.close()
    (cla, ta), (clb, tb) = outs
    assert [list(c) for c in cla] == [list(c) for c in clb] and len(ta) == len(tb)
    for (la, ia, ma), (lb, ib, mb) in zip(ta, tb):
        assert la == lb and ia == ib and np.array_equal(ma, mb)
