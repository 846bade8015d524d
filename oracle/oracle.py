"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the nlmc_b200 CUDA path.

Python driver for ``oracle/nlmc_oracle.c`` plus run-level restatements of the reference's four
``run()`` methods.  Nothing under ``nonlocal-monte-carlo_b200/`` imports this module; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs do.

Parity status: PINNED -- ``tests/test_oracle_vs_reference.py`` checks every function here
bit-for-bit against the live reference (when /root/reference is mounted) and
``tests/test_oracle_golden.py`` checks it against the committed golden vectors everywhere.

Random streams.  The reference draws from the global legacy ``np.random`` and from ``random``.
Its process pools are only reproducible with ``num_cores=1``; then there is one forked worker
whose generator state is a copy of the parent's state at the first ``submit``.
``fork_rng()`` reproduces exactly that.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import random
import subprocess

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i8p = np.ctypeslib.ndpointer(np.int8, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc, no other dependency)."""
    so = os.path.join(_HERE, "libnlmc_oracle.so")
    srcs = [os.path.join(_HERE, "nlmc_oracle.c"), os.path.join(_HERE, "npmath_host.c"),
            os.path.join(_HERE, "..", "nonlocal-monte-carlo_b200", "csrc", "nlmc_npmath.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.nlmc_oracle_mcmc.restype = C.c_int
        L.nlmc_oracle_mcmc.argtypes = [C.c_int, _i32p, _i32p, _f64p, _f64p, C.c_int, _f64p, _i32p, _f64p,
                                       _i8p, C.c_void_p, C.c_void_p, C.c_int]
        L.nlmc_oracle_energy.restype = C.c_int
        L.nlmc_oracle_energy.argtypes = [C.c_int, _i32p, _i32p, _f64p, _f64p, C.c_int, _i8p, _f64p]
        L.nlmc_oracle_disagreement_clusters.restype = C.c_int
        L.nlmc_oracle_disagreement_clusters.argtypes = [C.c_int, _i32p, _i32p, _f64p, _i8p, _i8p, _i32p]
        L.nlmc_oracle_mcmc_many.restype = C.c_int
        L.nlmc_oracle_mcmc_many.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _f64p, _f64p, C.c_int, _f64p,
                                            _i32p, _f64p, _i8p, C.c_int]
        L.nlmc_oracle_lbp_gather.restype = C.c_int
        L.nlmc_oracle_lbp_gather.argtypes = [C.c_int, _i32p, _i32p, _i32p, _f64p, _f64p, _f64p, _f64p]
        L.nlmc_oracle_lbp_colsum.restype = C.c_int
        L.nlmc_oracle_lbp_colsum.argtypes = [C.c_int, _i32p, _i32p, _f64p, _f64p]
        for name in ("nlmc_oracle_np_tanh", "nlmc_oracle_np_arctanh"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [_f64p, _f64p, C.c_long]
        _LIB = L
    return _LIB


def npmath_host(which: str, x) -> np.ndarray:
    """Host build of the product's restated np.tanh / np.arctanh (oracle/npmath_host.c); ``which`` is
    'tanh' or 'arctanh'.  Used by tests/test_npmath.py to pin the restatement against numpy without a GPU."""
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
    out = np.empty_like(x)
    getattr(lib(), "nlmc_oracle_np_" + which)(x, out, x.size)
    return out


# ----------------------------------------------------------------------------------------------
# instance helpers
# ----------------------------------------------------------------------------------------------
class Csr:
    """CSR exactly as ``scipy.sparse.csr_matrix(J)`` builds it (NMC/nmc.py:53)."""

    def __init__(self, J):
        A = sp.csr_matrix(J)
        self.n = A.shape[0]
        self.rp = np.ascontiguousarray(A.indptr, dtype=np.int32)
        self.ci = np.ascontiguousarray(A.indices, dtype=np.int32)
        self.val = np.ascontiguousarray(A.data, dtype=np.float64)
        self._rev = None

    def with_values(self, val):
        out = Csr.__new__(Csr)
        out.n, out.rp, out.ci, out._rev = self.n, self.rp, self.ci, self._rev
        out.val = np.ascontiguousarray(val, dtype=np.float64)
        return out

    @property
    def row_of(self):
        return np.repeat(np.arange(self.n, dtype=np.int32), np.diff(self.rp))

    @property
    def rev(self):
        """index of entry (j,i) for every entry (i,j); requires a symmetric pattern"""
        if self._rev is None:
            rows = self.row_of.astype(np.int64)
            cols = self.ci.astype(np.int64)
            key = rows * self.n + cols
            rkey = cols * self.n + rows
            order = np.argsort(key, kind="stable")
            pos = np.searchsorted(key[order], rkey)
            if np.any(pos >= len(key)) or np.any(key[order][np.minimum(pos, len(key) - 1)] != rkey):
                raise ValueError("J must have a symmetric sparsity pattern")
            self._rev = np.ascontiguousarray(order[pos], dtype=np.int32)
        return self._rev


def max_abs(J) -> float:
    if sp.issparse(J):
        return float(abs(J).max())
    return float(np.max(np.abs(J)))


def ea3d_pm_j(L: int, seed: int):
    """3D periodic +-J Edwards-Anderson instance of SURVEY.md 8(d): site i = x + L*(y + L*z),
    three forward bonds per site, +-1 equiprobable, h = 0.  Returned as scipy CSR."""
    rs = np.random.RandomState(seed)
    N = L ** 3
    idx = np.arange(N)
    x, y, z = idx % L, (idx // L) % L, idx // (L * L)
    nbr = [((x + 1) % L) + L * (y + L * z), x + L * (((y + 1) % L) + L * z), x + L * (y + L * ((z + 1) % L))]
    vals = rs.choice([-1.0, 1.0], size=(3, N))
    rows = np.concatenate([idx, idx, idx])
    cols = np.concatenate(nbr)
    v = vals.reshape(-1)
    A = sp.coo_matrix((np.concatenate([v, v]), (np.concatenate([rows, cols]), np.concatenate([cols, rows]))),
                      shape=(N, N)).tocsr()
    A.sum_duplicates()
    A.sort_indices()
    return A, np.zeros(N)


def random_pm_graph(N: int, p: float, seed: int):
    """Config C1: each pair i<j present with probability p, value +-1."""
    rs = np.random.RandomState(seed)
    iu = np.triu_indices(N, 1)
    keep = rs.rand(len(iu[0])) < p
    v = rs.choice([-1.0, 1.0], size=int(keep.sum()))
    J = np.zeros((N, N))
    J[iu[0][keep], iu[1][keep]] = v
    J += J.T
    return J, np.zeros(N)


def sk_gaussian(N: int, seed: int):
    """Config C3: J_ij ~ N(0,1)/sqrt(N), symmetric, zero diagonal."""
    rs = np.random.RandomState(seed)
    iu = np.triu_indices(N, 1)
    J = np.zeros((N, N))
    J[iu] = rs.randn(len(iu[0])) / math.sqrt(N)
    J += J.T
    return J, np.zeros(N)


def to_i8(m) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(m).reshape(-1), dtype=np.int8)


def fork_rng() -> np.random.RandomState:
    """Generator of a pool worker forked now: a copy of the global np.random state."""
    rs = np.random.RandomState()
    rs.set_state(np.random.get_state())
    return rs


def draw_sweeps(rng, n_sweeps: int, n: int):
    """Per sweep one permutation(N) then N rand() (NMC/nmc.py:71,87)."""
    perm = np.empty((n_sweeps, n), dtype=np.int32)
    u = np.empty((n_sweeps, n), dtype=np.float64)
    for s in range(n_sweeps):
        perm[s] = rng.permutation(n)
        u[s] = rng.rand(n)
    return perm, u


def tanh_lut(beta_run: np.ndarray, half: int) -> np.ndarray:
    """numpy's own tanh(beta*f) for integer fields f in [-half, half]; shape (n_sweeps, 2*half+1)."""
    f = np.arange(-half, half + 1, dtype=np.float64)
    return np.ascontiguousarray(np.tanh(np.asarray(beta_run, dtype=np.float64)[:, None] * f[None, :]))


# ----------------------------------------------------------------------------------------------
# element level
# ----------------------------------------------------------------------------------------------
def anneal_schedule(num_sweeps, beta, anneal=False, sweeps_per_beta=1, initial_beta=0.0):
    """beta_run of NMC/nmc.py:56-69."""
    if num_sweeps < 0:
        raise ValueError("negative dimensions are not allowed")
    run = np.zeros(num_sweeps)
    if not anneal:
        run[:] = beta
        return run
    num_betas = num_sweeps // sweeps_per_beta
    vals = np.linspace(initial_beta, beta, num_betas)
    idx = 0
    for jj in range(num_sweeps):
        if jj % sweeps_per_beta == 0 and idx < num_betas - 1:
            idx += 1
        run[jj] = vals[idx]
    return run


def mcmc(csr: Csr, h_eff, m_start, beta_run, rng=None, perm=None, u=None, use_lut=True):
    """Reference MCMC (NMC/nmc.py:28-91).  Returns M as int8 [n_sweeps][n] and the final state."""
    n = csr.n
    beta_run = np.ascontiguousarray(beta_run, dtype=np.float64)
    n_sweeps = len(beta_run)
    if perm is None:
        perm, u = draw_sweeps(rng if rng is not None else np.random, n_sweeps, n)
    m = to_i8(m_start).copy()
    M = np.empty((n_sweeps, n), dtype=np.int8)
    h_eff = np.ascontiguousarray(np.asarray(h_eff, dtype=np.float64).reshape(-1))
    lut, half = None, 0
    if use_lut and n_sweeps and np.all(csr.val == np.round(csr.val)):
        half = int(np.bincount(csr.row_of, weights=np.abs(csr.val), minlength=n).max()) if csr.rp[-1] else 0
        lut_arr = tanh_lut(beta_run, half)
        lut = lut_arr.ctypes.data
    lib().nlmc_oracle_mcmc(n, csr.rp, csr.ci, csr.val, h_eff, n_sweeps, beta_run,
                           np.ascontiguousarray(perm, dtype=np.int32), np.ascontiguousarray(u, dtype=np.float64),
                           m, M.ctypes.data, lut, half)
    return M, m


def energy(csr: Csr, h, M):
    """E = -(m^T J m/2 + m^T h) per recorded state; M int8 [n_cols][n]."""
    M = np.ascontiguousarray(M, dtype=np.int8).reshape(-1, csr.n)
    E = np.empty(M.shape[0], dtype=np.float64)
    lib().nlmc_oracle_energy(csr.n, csr.rp, csr.ci, csr.val,
                             np.ascontiguousarray(np.asarray(h, dtype=np.float64).reshape(-1)), M.shape[0], M, E)
    return E


def disagreement_clusters(csr: Csr, s1, s2):
    """labels (cluster ordinal in min-site order, -1 = agree) and the cluster count (apt_ICM.py:116-143)."""
    labels = np.empty(csr.n, dtype=np.int32)
    k = lib().nlmc_oracle_disagreement_clusters(csr.n, csr.rp, csr.ci, csr.val, to_i8(s1), to_i8(s2), labels)
    return labels, k


# ----------------------------------------------------------------------------------------------
# LBP backbone search (NMC/nmc.py:93-318)
# ----------------------------------------------------------------------------------------------
def _pairwise_rowsum_abs(csr: Csr):
    """np.sum(np.abs(J), axis=1) of the dense matrix (NMC/nmc.py:353): numpy pairwise order."""
    out = np.zeros(csr.n)
    a = np.abs(csr.val)
    exact = np.all(a == np.round(a))
    for i in range(csr.n):
        b, e = csr.rp[i], csr.rp[i + 1]
        if exact:
            out[i] = a[b:e].sum()
        else:
            dense = np.zeros(csr.n)
            dense[csr.ci[b:e]] = a[b:e]
            out[i] = np.sum(dense)
    return out


def find_clusters(csr: Csr, marg, thr_init, thr_cut, thr_step):
    """Backbone seeds and growth; semantics of find_clusters (NMC/nmc.py:257-318)."""
    marg = np.asarray(marg)
    seeds = np.where(np.abs(marg) >= thr_init)[0]
    seed_set = set(int(s) for s in seeds)
    taken: set[int] = set()
    clusters: list[list[int]] = []

    def nbrs(i):
        b, e = csr.rp[i], csr.rp[i + 1]
        return sorted(set(int(j) for j, v in zip(csr.ci[b:e], csr.val[b:e]) if v != 0))

    for s in seeds:
        s = int(s)
        if s in taken:
            continue
        free = [j for j in nbrs(s) if j not in taken]
        cl = [s] + [j for j in free if j in seed_set]
        clusters.append(cl)
        taken.update(cl)
    cur = thr_init - thr_step
    while cur > thr_cut:
        for i, cl in enumerate(clusters):
            cand = sorted(set(j for k in cl for j in nbrs(k)) - taken)
            add = [j for j in cand if abs(marg[j]) >= cur]
            clusters[i] = cl + add
            taken.update(add)
        cur -= thr_step
    return [np.array(c, dtype=int) for c in clusters]


def _atanh_saturated(x):
    """NMC/nmc.py:230-255, with numpy's own tanh/arctanh (see nlmc_oracle_lbp_gather)."""
    e = np.finfo(float).eps
    return np.arctanh(np.clip(x, np.tanh(-19.06) + e, np.tanh(19.06) - e))


def lbp(csr: Csr, hl, beta, u, hm, tot, tol, max_iter):
    """LoopyBeliefPropagation (NMC/nmc.py:168-228) on the stored entries of J.
    u, hm (per entry) and tot (off-edge value of each h_msgs row) are updated in place.
    Returns (marginal, iteration) with `iteration` the reference's loop variable on exit."""
    n = csr.n
    has_offedge = np.array([np.count_nonzero(csr.ci[csr.rp[i]:csr.rp[i + 1]] != i) < n - 1 for i in range(n)])
    tj = np.tanh(beta * csr.val)
    iteration = max_iter - 1
    for iteration in range(max_iter):
        u_old, hm_old, tot_old = u.copy(), hm.copy(), tot.copy()
        lib().nlmc_oracle_lbp_gather(n, csr.rp, csr.ci, csr.rev, hl, u_old, hm, tot)
        u[:] = (1 / beta) * _atanh_saturated(tj * np.tanh(beta * hm))
        with np.errstate(invalid="ignore", divide="ignore"):
            u_change = np.max(np.abs(u - u_old)) / np.max(np.abs(u) + np.abs(u_old))
            dh = np.max(np.abs(hm - hm_old), initial=0.0)
            sh = np.max(np.abs(hm) + np.abs(hm_old), initial=0.0)
            if has_offedge.any():
                dh = max(dh, np.max(np.abs(tot - tot_old)[has_offedge]))
                sh = max(sh, np.max((np.abs(tot) + np.abs(tot_old))[has_offedge]))
            h_change = dh / sh
        if u_change < tol and h_change < tol:
            break
    acc = np.empty(n)
    lib().nlmc_oracle_lbp_colsum(n, csr.rp, csr.rev, u, acc)
    return np.tanh(beta * (hl + acc)), iteration


def lbp_byproducts(csr: Csr, beta, hm, tot, marginal):
    """correlations, h_tilde, J_tilde of one LBP call, dense like the reference's (NMC/nmc.py:217-226).
    Off the stored entries h_msgs[i, j] = tot[i] (j != i) and the diagonal of h_msgs is 0 (nmc.py:203)."""
    n = csr.n
    Jd = np.zeros((n, n))
    Jd[csr.row_of, csr.ci] = csr.val
    H = np.repeat(np.asarray(tot, dtype=np.float64)[:, None], n, axis=1)
    np.fill_diagonal(H, 0.0)
    H[csr.row_of, csr.ci] = hm
    tJ, tH = np.tanh(beta * Jd), np.tanh(beta * H)
    corr = (tJ + tH * tH.T) / (1 + tJ * tH * tH.T + 1e-10)
    corr = corr - np.diag(np.diag(corr))
    return corr, (1 / beta) * _atanh_saturated(marginal), (1 / beta) * _atanh_saturated(corr)


def lbp_dense(J, h, beta, h_msgs, u_msgs, tol, max_iter):
    """LoopyBeliefPropagation with the reference's dense arguments and return tuple (NMC/nmc.py:168-228), for
    message matrices of the form the reference itself produces (u_msgs zero and h_msgs row-constant off J's entries)."""
    csr = Csr(J)
    n = csr.n
    r, c = csr.row_of, csr.ci
    h_msgs, u_msgs = np.asarray(h_msgs, dtype=np.float64), np.asarray(u_msgs, dtype=np.float64)
    off = np.ones((n, n), dtype=bool)
    off[r, c] = False
    np.fill_diagonal(off, False)
    assert not np.any(u_msgs[off]), "u_msgs must vanish off the entries of J"
    tot = np.where(off.any(axis=1), h_msgs[np.arange(n), np.argmax(off, axis=1)], 0.0)
    assert np.all(h_msgs[off] == np.repeat(tot[:, None], n, axis=1)[off]), "h_msgs rows must be constant off J"
    u = np.ascontiguousarray(u_msgs[r, c])
    hm = np.ascontiguousarray(h_msgs[r, c])
    hl = np.ascontiguousarray(np.asarray(h, dtype=np.float64).reshape(-1))
    marg, it = lbp(csr, hl, beta, u, hm, tot, tol, max_iter)
    corr, ht, jt = lbp_byproducts(csr, beta, hm, tot, marg)
    H = np.repeat(tot[:, None], n, axis=1)
    np.fill_diagonal(H, 0.0)
    H[r, c] = hm
    U = np.zeros((n, n))
    U[r, c] = u
    return marg, corr, ht, jt, it, H, U


def lbp_convexified(csr: Csr, h, m_star, epsilon, lambda_start, lambda_end, factor, tol, max_iter,
                    thr_init, thr_cut, beta):
    """lambda-annealed LBP (NMC/nmc.py:93-166).  Returns (clusters, marginal, n_lambda_steps)."""
    h = np.asarray(h, dtype=np.float64).reshape(-1)
    m_star = np.asarray(m_star, dtype=np.float64).reshape(-1)
    u = np.ascontiguousarray(csr.val * m_star[csr.ci])  # J * m_star.reshape(1,-1)  nmc.py:129
    hm = np.zeros_like(u)
    tot = np.zeros(csr.n)
    lam = lambda_start
    marg = np.zeros(csr.n)
    prev = None
    steps = 0
    while lam >= lambda_end:
        hl = np.ascontiguousarray(h + lam * m_star * epsilon)
        marg, it = lbp(csr, hl, beta, u, hm, tot, tol, max_iter)
        steps += 1
        if it == max_iter - 1 and lam == lambda_start:
            raise ValueError('LBP diverged at initial lambda, please try a larger lambda_start or increase '
                             'max_iterations or beta')
        elif it == max_iter - 1:
            lambda_end = lam
            marg = prev.copy()
        else:
            prev = marg.copy()
        lam = lam * factor
        if round(lam, 6) == 0:
            break
    return find_clusters(csr, marg, thr_init, thr_cut, 0.01), marg, steps


# ----------------------------------------------------------------------------------------------
# NMC_subroutine, both variants (NMC/nmc.py:320-440 ; NPT/npt.py:357-477)
# ----------------------------------------------------------------------------------------------
def nmc_subroutine(csr: Csr, h, m_star, num_cycles, phase_sweeps, full_update_frequency, M_skip, global_beta,
                   temp_x, lambda_start, lambda_end, factor, thr_init, thr_cut, max_iter, tol,
                   variant: str, rng=None, all_clusters=None):
    rng = rng if rng is not None else np.random
    n = csr.n
    h = np.asarray(h, dtype=np.float64).reshape(-1)
    eps = np.abs(h) + _pairwise_rowsum_abs(csr)
    m_init = np.asarray(m_star, dtype=np.float64).reshape(-1).copy()
    m_star = m_init.copy()
    cap = phase_sweeps * num_cycles * 3 // M_skip
    M_overall = np.zeros((n, cap))
    E_overall = np.zeros(cap)
    idx = 0
    beta_run = np.full(phase_sweeps, float(global_beta))
    rows = csr.row_of

    def backbone(ms):
        cl, _, _ = lbp_convexified(csr, h, ms, eps, lambda_start, lambda_end, factor, tol, max_iter,
                                   thr_init, thr_cut, global_beta)
        return np.concatenate(cl).astype(int) if cl else np.array([], dtype=int)

    def record(Mi8):
        nonlocal idx, m_init
        E = energy(csr, h, Mi8)
        Mf = Mi8.T.astype(np.float64)
        w = phase_sweeps // M_skip
        M_overall[:, idx:idx + w] = Mf[:, ::M_skip]
        E_overall[idx:idx + w] = E[::M_skip]
        idx += w
        m_init = Mf[:, int(np.argmin(E))].copy()
        return E

    all_cl = None if all_clusters is None else np.asarray(all_clusters, dtype=int)
    if variant == "npt" and all_clusters is None:
        all_cl = backbone(m_star)
    for cycle in range(num_cycles):
        if variant == "nmc" and all_clusters is None:
            all_cl = backbone(m_star)
        non_cl = np.setdiff1d(np.arange(n), all_cl)
        in_cl = np.zeros(n, dtype=bool)
        in_cl[all_cl] = True
        # phase C: backbone rows /temp_x, everything else frozen by +-1e4 (nmc.py:377-385)
        val_c = np.where(in_cl[rows], csr.val / temp_x, csr.val)
        h_c = h.copy()
        h_c[all_cl] /= temp_x
        h_c[non_cl] = m_init[non_cl] * 10000
        M, _ = mcmc(csr.with_values(val_c), h_c, m_init, beta_run, rng=rng)
        record(M)
        # phase NC: backbone frozen (nmc.py:398-406)
        h_nc = h.copy()
        h_nc[all_cl] = m_init[all_cl] * 10000
        M, _ = mcmc(csr, h_nc, m_init, beta_run, rng=rng)
        record(M)
        # phase ALL (nmc.py:419-433)
        if cycle % full_update_frequency == 0:
            M, _ = mcmc(csr, h, m_init, beta_run, rng=rng)
            record(M)
            if variant == "nmc":
                m_star = m_init.copy()
    M_overall = M_overall[:, :idx]
    E_overall = E_overall[:idx]
    return M_overall, E_overall, np.min(E_overall), all_cl


# ----------------------------------------------------------------------------------------------
# run() restatements
# ----------------------------------------------------------------------------------------------
def nmc_run(J, h, num_sweeps_initial=10000, num_sweeps_per_NMC_phase=10000, num_NMC_cycles=10,
            full_update_frequency=1, M_skip=1, temp_x=20, global_beta=2.5, lambda_start=0.5, lambda_end=0.01,
            lambda_reduction_factor=0.9, threshold_initial=0.999999, threshold_cutoff=0.99999,
            max_iterations=100, tolerance=np.finfo(float).eps):
    """NMC.run (NMC/nmc.py:442-520)."""
    norm = max_abs(J)
    csr = Csr(J / norm)
    hn = np.asarray(h, dtype=np.float64).reshape(-1) / norm
    n = csr.n
    m_init = np.sign(2 * np.random.rand(n) - 1)
    M, _ = mcmc(csr, hn, m_init, anneal_schedule(num_sweeps_initial, global_beta, True, 1, 0))
    E = energy(csr, hn, M)
    m_star = M[int(np.argmin(E))].astype(np.float64)
    M_o, E_o, mn, _ = nmc_subroutine(csr, hn, m_star, num_NMC_cycles, num_sweeps_per_NMC_phase,
                                     full_update_frequency, M_skip, global_beta, temp_x, lambda_start,
                                     lambda_end, lambda_reduction_factor, threshold_initial, threshold_cutoff,
                                     max_iterations, tolerance, "nmc")
    return M_o, E_o, mn


def select_non_overlapping_pairs(all_pairs, k):
    """NPT/npt.py:514-533 (random.randint from the global `random`)."""
    avail = list(all_pairs)
    chosen = []
    for _ in range(k):
        if not avail:
            raise ValueError("Cannot find non-overlapping pairs.")
        pair = avail[random.randint(0, len(avail) - 1)]
        chosen.append(pair)
        avail = [p for p in avail if p[0] not in pair and p[1] not in pair]
    return chosen


def npt_run(J, h, beta_list, num_replicas, doNMC, num_sweeps_MCMC=1000, num_sweeps_read=1000,
            num_swap_attempts=100, num_swapping_pairs=1, num_cycles=10, full_update_frequency=1, M_skip=1,
            temp_x=20, global_beta=2.5, lambda_start=0.5, lambda_end=0.01, lambda_reduction_factor=0.9,
            threshold_initial=0.999999, threshold_cutoff=0.99999, max_iterations=100,
            tolerance=np.finfo(float).eps):
    """NPT.run with num_cores=1 (NPT/npt.py:535-700)."""
    R = num_replicas
    spm = num_sweeps_MCMC // num_swap_attempts
    spr = num_sweeps_read // num_swap_attempts
    phase = int(np.ceil(num_sweeps_MCMC / num_swap_attempts / 3 / num_cycles))
    norm = max_abs(J)
    csr = Csr(J / norm)
    hn = np.asarray(h, dtype=np.float64).reshape(-1) / norm
    if len(doNMC) != R:
        raise ValueError("The length of doNMC does not match the number of replicas.")
    n = csr.n
    pairs_all = [(i, i + 1) for i in range(1, R)]
    M = np.zeros((R * n, spm))
    m_start = np.sign(2 * np.random.rand(R * n, 1) - 1)
    worker = None
    for _ in range(num_swap_attempts):
        if worker is None:
            worker = fork_rng()
        for r in range(R):
            ms = m_start[r * n:(r + 1) * n].reshape(-1)
            if not doNMC[r]:
                Mi8, _ = mcmc(csr, hn, ms, np.full(spm, float(beta_list[r])), rng=worker)
                Mr = Mi8.T.astype(np.float64)
            else:
                Mr, _, _, _ = nmc_subroutine(csr, hn, ms, num_cycles, phase, full_update_frequency, M_skip,
                                             global_beta, temp_x, lambda_start, lambda_end,
                                             lambda_reduction_factor, threshold_initial, threshold_cutoff,
                                             max_iterations, tolerance, "npt", rng=worker)
            M[r * n:(r + 1) * n, :] = Mr[:, -spm:]
        m_start = M[:, -1].copy().reshape(-1, 1)
        last = M[:, -1]
        for sel, nxt in select_non_overlapping_pairs(pairs_all, num_swapping_pairs):
            m_sel = last[(sel - 1) * n:sel * n].copy()
            m_nxt = last[(nxt - 1) * n:nxt * n].copy()
            E_sel, E_nxt = energy(csr, hn, np.stack([to_i8(m_sel), to_i8(m_nxt)]))
            dE = E_nxt - E_sel
            dB = beta_list[nxt - 1] - beta_list[sel - 1]
            if np.random.rand() < min(1, np.exp(dB * dE)):
                m_start[(sel - 1) * n:sel * n] = m_nxt.reshape(-1, 1)
                m_start[(nxt - 1) * n:nxt * n] = m_sel.reshape(-1, 1)
    Energy = np.zeros(R)
    for r in range(R):
        cols = M[r * n:(r + 1) * n, :spr]
        Energy[r] = np.min(energy(csr, hn, to_i8(cols.T).reshape(-1, n)))
    return M, Energy


def apt_preprocessor_run(J, h, num_sweeps_MCMC=1000, num_sweeps_read=1000, num_rng=100, beta_start=0.5,
                         alpha=1.25, sigma_E_val=1000, beta_max=30, write_files=False):
    """APT_preprocessor.run with num_cores=1 (NPT/apt_preprocessor.py:115-204)."""
    norm = max_abs(J)
    csr = Csr(J / norm)
    hn = np.asarray(h, dtype=np.float64).reshape(-1) / norm
    n = csr.n
    beta = [beta_start]
    sigma_E = sigma_E_val
    sigma_min = 0.5 * np.min(np.abs(csr.val[csr.val != 0]))
    sigma = []
    saved = np.zeros((num_rng, n))
    it = 1
    while sigma_E > sigma_min:
        if it != 1:
            beta.append(beta[-1] + alpha / sigma_E)
        Energy = np.zeros((num_rng, num_sweeps_read))
        worker = None
        for j in range(num_rng):
            if it == 1:
                ms = np.sign(2. * np.random.rand(n, 1) - 1).reshape(-1)
            else:
                ms = saved[j].copy()
            if worker is None:
                worker = fork_rng()  # new pool every iteration (apt_preprocessor.py:160)
            Mi8, _ = mcmc(csr, hn, ms, anneal_schedule(num_sweeps_MCMC, beta[-1]), rng=worker)
            tail = Mi8[-num_sweeps_read:]
            Energy[j, :] = energy(csr, hn, tail)
            saved[j, :] = tail[-1]
        sigma_E = np.mean(np.std(Energy, axis=1))
        if beta[-1] > beta_max:
            break
        sigma.append(sigma_E)
        it += 1
    return beta, sigma


def apt_icm_run(J, h, beta_list, num_replicas, num_sweeps_MCMC=1000, num_sweeps_read=1000,
                num_swap_attempts=100, num_swapping_pairs=1):
    """APT_ICM.run (NPT/apt_ICM.py:145-305); J is NOT normalised inside run."""
    S = 10  # num_subreplicas, apt_ICM.py:177
    R = num_replicas
    spm = num_sweeps_MCMC // num_swap_attempts
    spr = num_sweeps_read // num_swap_attempts
    csr = Csr(J)
    hn = np.asarray(h, dtype=np.float64).reshape(-1)
    n = csr.n
    pairs_all = [(i, i + 1) for i in range(1, R)]
    M = np.zeros((n * R, spm * S))
    m_start = np.sign(2 * np.random.rand(n * R, S) - 1)
    for _ in range(int(num_swap_attempts)):
        for r in range(R):
            for s in range(S):
                Mi8, last = mcmc(csr, hn, m_start[r * n:(r + 1) * n, s], np.full(spm, float(beta_list[r])))
                M[r * n:(r + 1) * n, s * spm:(s + 1) * spm] = Mi8.T
                m_start[r * n:(r + 1) * n, s] = last
        for r in range(R):
            shuffled = np.random.permutation(S)
            for p in range(S // 2):
                a, b = shuffled[2 * p], shuffled[2 * p + 1]
                s1 = M[r * n:(r + 1) * n, a * spm].copy()
                s2 = M[r * n:(r + 1) * n, b * spm].copy()
                labels, k = disagreement_clusters(csr, s1, s2)
                if k:
                    pick = np.random.randint(k)
                    members = labels == pick
                    if int(members.sum()) > n // 2:  # Katzgraber rule, apt_ICM.py:236-237
                        s1 = -s1
                    else:
                        s1[members], s2[members] = s2[members].copy(), s1[members].copy()
                    M[r * n:(r + 1) * n, a * spm] = s1
                    M[r * n:(r + 1) * n, b * spm] = s2
        selected = select_non_overlapping_pairs(pairs_all, num_swapping_pairs)
        for s in range(S):
            last = M[:, (s + 1) * spm - 1]
            for sel, nxt in selected:
                m_sel = last[(sel - 1) * n:sel * n].copy()
                m_nxt = last[(nxt - 1) * n:nxt * n].copy()
                E_sel, E_nxt = energy(csr, hn, np.stack([to_i8(m_sel), to_i8(m_nxt)]))
                dE = E_nxt - E_sel
                dB = beta_list[nxt - 1] - beta_list[sel - 1]
                if np.random.rand() < min(1, np.exp(dB * dE)):
                    m_start[(sel - 1) * n:sel * n, s] = m_nxt
                    m_start[(nxt - 1) * n:nxt * n, s] = m_sel
    Energy = np.zeros(R)
    for r in range(R):
        cols = M[r * n:(r + 1) * n, :spr]
        Energy[r] = np.min(energy(csr, hn, to_i8(cols.T).reshape(-1, n)))
    return M, Energy
