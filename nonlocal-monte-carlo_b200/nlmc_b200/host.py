"""Host-side logic shared by the drop-in classes: instance preparation, the reference's random
streams (exact-replay mode), schedules and replica-pair selection.  No arithmetic of the hot path
happens here -- sweeps, energies, LBP, cluster search and swaps run in libnlmc_b200.so.
"""
from __future__ import annotations

import random

import numpy as np
import scipy.sparse as sp

from . import _lib


def max_abs(J) -> float:
    """np.max(np.abs(J)) for dense or scipy.sparse J (NMC/nmc.py:474)."""
    if sp.issparse(J):
        if not J.nnz:
            return 0.0
        data = getattr(J, "data", None)   # CSR / CSC / COO / BSR keep their entries in .data: no |J| copy of the matrix
        if isinstance(data, np.ndarray) and J.format in ("csr", "csc", "coo", "bsr"):
            return float(max(data.max(), -data.min(), 0.0))
        return float(abs(J).max())
    return float(np.max(np.abs(J)))


def normalised(J, h, norm_factor: float):
    """J / norm_factor, h / norm_factor as new objects (the reference rebinds self.J / self.h, NPT/npt.py:588-590);
    a factor of exactly 1 leaves every value unchanged, so the division is skipped."""
    if norm_factor == 1.0 and getattr(J, "dtype", None) == np.float64 and getattr(h, "dtype", None) == np.float64:
        return J, h
    return J / norm_factor, h / norm_factor


class Problem:
    """Normalised instance on the device plus what the host needs to drive it."""

    def __init__(self, J, h, device: int = 0):
        A = sp.csr_matrix(J)  # entry order = scipy's, which the reference's J.dot(m) uses (NMC/nmc.py:53,86)
        self.n = A.shape[0]
        if A.shape[0] != A.shape[1]:
            raise ValueError("J must be square")
        self.rp = np.ascontiguousarray(A.indptr, dtype=np.int32)
        self.ci = np.ascontiguousarray(A.indices, dtype=np.int32)
        self.val = np.ascontiguousarray(A.data, dtype=np.float64)
        self.h = np.ascontiguousarray(np.asarray(h, dtype=np.float64).reshape(-1))
        self.inst = _lib.Instance(self.rp, self.ci, self.val, self.h, device)
        self.is_integer = self.inst.is_integer
        self._row_of = None
        self._lut_half = None

    @property
    def row_of(self) -> np.ndarray:
        """Row index of every stored entry (built on first use: the production paths never need it)."""
        if self._row_of is None:
            self._row_of = np.repeat(np.arange(self.n, dtype=np.int32), np.diff(self.rp))
        return self._row_of

    @property
    def lut_half(self) -> int:
        """Largest |integer field| a site can see: half-width of the tanh table of the replay kernel."""
        if self._lut_half is None:
            if self.is_integer and len(self.val):
                self._lut_half = int(np.bincount(self.row_of, weights=np.abs(self.val), minlength=self.n).max())
            else:
                self._lut_half = 0
        return self._lut_half

    def lbp_instance(self):
        """Instance for the belief-propagation kernel (K5), which needs rows sorted by column (numpy sums the reference's
        dense rows and columns in index order): the instance itself when scipy delivered sorted rows -- always the case
        for the dense J the reference's LBP requires -- otherwise a sorted twin, built once."""
        if getattr(self, "_lbp_inst", None) is None:
            A = sp.csr_matrix((self.val, self.ci, self.rp), shape=(self.n, self.n))
            if A.has_sorted_indices:
                self._lbp_inst = self.inst
            else:
                A = A.copy()
                A.sort_indices()
                self._lbp_inst = _lib.Instance(A.indptr, A.indices, A.data, self.h, self.inst.device)
        return self._lbp_inst

    def tanh_lut(self, beta_sched: np.ndarray):
        """tanh(beta*f) for every integer field f, computed with numpy's own tanh so that decisions on
        +-J instances are bit-equal to the reference's np.tanh(beta_run[jj] * x[kk]) (NMC/nmc.py:87).
        beta_sched [G][S] -> [G][S][2*half+1] or None when the instance is not integer-valued."""
        if not self.is_integer:
            return None
        f = np.arange(-self.lut_half, self.lut_half + 1, dtype=np.float64)
        return np.ascontiguousarray(np.tanh(np.asarray(beta_sched, dtype=np.float64)[..., None] * f))

    def neighbours(self, i: int):
        b, e = self.rp[i], self.rp[i + 1]
        return self.ci[b:e][self.val[b:e] != 0]


def as_spins_i8(m) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(m).reshape(-1), dtype=np.int8)


def fork_rng() -> np.random.RandomState:
    """State of a ProcessPoolExecutor worker forked now (reference with num_cores=1): a copy of the
    global np.random generator (NPT/npt.py:616, NPT/apt_preprocessor.py:160)."""
    rs = np.random.RandomState()
    rs.set_state(np.random.get_state())
    return rs


def draw_sweeps(rng, n_sweeps: int, n: int):
    """The reference's draws for n_sweeps sweeps: per sweep np.random.permutation(N), then N scalar
    np.random.rand() in visit order (NMC/nmc.py:71,87); N scalar draws equal one rand(N)."""
    perm = np.empty((n_sweeps, n), dtype=np.int32)
    u = np.empty((n_sweeps, n), dtype=np.float64)
    for s in range(n_sweeps):
        perm[s] = rng.permutation(n)
        u[s] = rng.rand(n)
    return perm, u


def beta_schedule(num_sweeps: int, beta: float, anneal: bool = False, sweeps_per_beta: int = 1,
                  initial_beta: float = 0.0) -> np.ndarray:
    """beta_run[jj] of MCMC (NMC/nmc.py:56-69)."""
    if num_sweeps < 0:
        raise ValueError("negative dimensions are not allowed")  # np.zeros((N, num_sweeps)) in the reference
    run = np.full(num_sweeps, float(beta))
    if anneal:
        num_betas = num_sweeps // sweeps_per_beta
        vals = np.linspace(initial_beta, beta, num_betas)
        idx = 0
        for jj in range(num_sweeps):
            if jj % sweeps_per_beta == 0 and idx < num_betas - 1:
                idx += 1
            run[jj] = vals[idx]
    return run


def select_non_overlapping_pairs(all_pairs, num_swapping_pairs: int):
    """NPT/npt.py:514-533 == NPT/apt_ICM.py:95-114; draws from the global `random` like the reference."""
    available = list(all_pairs)
    selected = []
    for _ in range(num_swapping_pairs):
        if not available:
            raise ValueError("Cannot find non-overlapping pairs.")
        pair = available[random.randint(0, len(available) - 1)]
        selected.append(pair)
        available = [p for p in available if p[0] not in pair and p[1] not in pair]
    return selected


def replay_chains(prob: Problem, reps: "_lib.Replicas", m_start, beta_sched, rng, record_from=0):
    """Run one batch of exact-replay chains: chain g starts from m_start[g], uses beta_sched[g][:] and
    consumes `rng` exactly as the reference's MCMC would when the chains are executed one after
    the other (g = 0, 1, ...).  Returns (M int8 [G][S-record_from][n], E [G][S])."""
    G = reps.R
    beta_sched = np.asarray(beta_sched, dtype=np.float64).reshape(G, -1)
    S = beta_sched.shape[1]
    n = prob.n
    perm = np.empty((G, S, n), dtype=np.int32)
    u = np.empty((G, S, n), dtype=np.float64)
    for g in range(G):
        perm[g], u[g] = draw_sweeps(rng, S, n)
    reps.set_spins(np.asarray(m_start).reshape(G, n))
    lut = prob.tanh_lut(beta_sched)
    return reps.sweep_replay(perm, u, beta_sched, lut, prob.lut_half, record_from=record_from)
