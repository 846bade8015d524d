"""Element-level public methods of the reference classes on the CUDA path.

The reference's classes expose the pieces of the hot path as public methods besides ``run()``:
``MCMC``, ``LoopyBeliefPropagation``, ``LBP_convexified``, ``atanh_saturated``, ``find_clusters``,
``NMC_subroutine`` (NMC/nmc.py:28-440, NPT/npt.py:47-477), ``MCMC_task`` / ``NMC_task``
(NPT/npt.py:112-127,479-512; NPT/apt_preprocessor.py:76-113) and ``replica_energy``
(NPT/npt.py:31-45, NPT/apt_ICM.py:36-50).  The mixins below give the drop-in classes the same methods with
the same arguments and return values; the arithmetic runs in libnlmc_b200.so (K1 sweeps, K4 energies,
K5 belief propagation), the host only moves data between the reference's dense layouts and the device's.

In ``mode="replay"`` the methods consume the global ``np.random`` stream exactly like the reference's, so
a seeded call returns the reference's arrays bit for bit (sweeps, energies) or to floating-point
tolerance (belief propagation: CUDA's tanh/atanh differ from numpy's in the last place).
"""
from __future__ import annotations

import zlib
from collections import OrderedDict, defaultdict

import numpy as np
import scipy.sparse as sp

from . import _lib, host
from . import nmc_core


def _fingerprint(J, h) -> tuple:
    """Content key of (J, h): the methods take J and h per call (the reference's NMC phases pass modified
    copies), so the device instance is cached by content, not by object identity."""
    if sp.issparse(J):
        A = J.tocsr()
        parts = (A.indptr, A.indices, A.data)
    else:
        parts = (np.ascontiguousarray(J),)
    crc = 0
    for p in parts + (np.ascontiguousarray(np.asarray(h, dtype=np.float64)),):
        crc = zlib.crc32(memoryview(np.ascontiguousarray(p)).cast("B"), crc)
    return (J.shape, crc)


def _check_hash_table(hash_table, use_hash_table, any_attempt: bool):
    """The reference validates the table at the first attempt of a sweep (NMC/nmc.py:74-76); the table itself is
    a CPU memoisation with no effect on results and is not used here."""
    if use_hash_table and any_attempt:
        try:
            from cachetools import LRUCache
            ok = isinstance(hash_table, LRUCache)
        except ImportError:  # cachetools absent: accept anything that is called LRUCache
            ok = type(hash_table).__name__ == "LRUCache"
        if not ok:
            raise ValueError("hash_table must be an instance of cachetools.LRUCache")


class _ProblemCache:
    """Small LRU of device instances keyed by the content of (J, h)."""

    def _problem_for(self, J, h) -> host.Problem:
        cache = self.__dict__.setdefault("_problems", OrderedDict())
        key = _fingerprint(J, h)
        prob = cache.get(key)
        if prob is None:
            prob = host.Problem(J, h, self.device)
            cache[key] = prob
            while len(cache) > 4:
                cache.popitem(last=False)
        else:
            cache.move_to_end(key)
        return prob


class SweepMethods(_ProblemCache):
    """``MCMC`` with the NMC/NPT signature (J and h passed per call)."""

    def _mcmc(self, num_sweeps, m_start, beta, J, h, anneal, sweeps_per_beta, initial_beta, hash_table,
              use_hash_table):
        N = J.shape[0]
        m0 = np.asarray(m_start).copy().reshape(-1)
        if len(m0) != N:
            raise ValueError(f"m_start has {len(m0)} entries, J has {N} rows")
        if not np.all(np.isin(m0, (-1, 0, 1))):
            raise ValueError("m_start must hold spins -1/+1 (0 allowed, as np.sign can produce it)")
        sched = host.beta_schedule(num_sweeps, beta, anneal, sweeps_per_beta, initial_beta)
        _check_hash_table(hash_table, use_hash_table, num_sweeps > 0 and N > 0)
        if num_sweeps == 0:
            return np.zeros((N, 0))
        prob = self._problem_for(J, h)
        if self.mode == "replay":
            reps = _lib.Replicas(prob.inst, 1)
            try:
                Mi8, _ = host.replay_chains(prob, reps, m0[None, :], sched[None, :], np.random)
            finally:
                reps.close()
            return Mi8[0].T.astype(np.float64)
        from .production import _generic_engine, _seed_from_numpy
        eng = _generic_engine(prob, [float(beta)], _seed_from_numpy())
        try:
            eng.set_spins(m0[None, :].astype(np.int8))
            if isinstance(eng, _lib.Col):
                states, _ = eng.sweep_record(num_sweeps, beta_sched=sched[:, None], want_energies=False)
                M = states[:, 0, :].T.astype(np.float64)
            else:
                M = np.zeros((N, num_sweeps))
                for jj in range(num_sweeps):
                    eng.set_betas([sched[jj]])
                    eng.sweep(1)
                    M[:, jj] = eng.get_spins()[0]
        finally:
            eng.close()
        return M

    def MCMC(self, num_sweeps, m_start, beta, J, h, anneal=False, sweeps_per_beta=1, initial_beta=0,
             hash_table=None, use_hash_table=False):
        """Heat-bath sweeps in random-permutation order (NMC/nmc.py:28-91 == NPT/npt.py:47-110).
        Returns M (N, num_sweeps) float64, column jj = the state after sweep jj."""
        return self._mcmc(num_sweeps, m_start, beta, J, h, anneal, sweeps_per_beta, initial_beta, hash_table,
                          use_hash_table)


class SweepMethodsFixedInstance(_ProblemCache):
    """``MCMC`` with the APT_preprocessor/APT_ICM signature (uses self.J, self.h; no annealing)."""

    def MCMC(self, num_sweeps, m_start, beta, hash_table=None, use_hash_table=False):
        """NPT/apt_preprocessor.py:33-74 == NPT/apt_ICM.py:52-93."""
        return SweepMethods._mcmc(self, num_sweeps, m_start, beta, self.J, self.h, False, 1, 0, hash_table,
                                  use_hash_table)


class EnergyMethods(_ProblemCache):
    def replica_energy(self, M, num_sweeps):
        """(min energy, energies) of the first num_sweeps columns of M (N x S), kernel K4
        (NPT/npt.py:31-45 == NPT/apt_ICM.py:36-50)."""
        prob = self._problem_for(self.J, self.h)
        cols = np.asarray(M)[:, :num_sweeps]
        if cols.shape[1] != num_sweeps:
            raise IndexError(f"index {cols.shape[1]} is out of bounds for axis 1 with size {cols.shape[1]}")
        EE1 = prob.inst.energy_states(np.ascontiguousarray(cols.T, dtype=np.int8))
        return np.min(EE1), EE1


class LbpMethods(_ProblemCache):
    """Belief-propagation backbone search and the NMC cycle (NMC/nmc.py:93-440, NPT/npt.py:129-477)."""

    _nmc_variant = "nmc"

    def atanh_saturated(self, x):
        """np.arctanh with the argument clipped to +-(1 - eps) (NMC/nmc.py:230-255), evaluated by the same device
        function the LBP kernel uses (bit-equal to np.arctanh)."""
        eps = np.finfo(float).eps
        xc = np.clip(x, -1.0 + eps, 1.0 - eps)              # tanh(+-19.06) == +-1.0 exactly
        out = _lib.np_arctanh(xc)
        return out if np.ndim(x) else np.float64(out.reshape(-1)[0])

    def find_clusters(self, magnetizations, threshold_initial, threshold_cutoff, threshold_step):
        """NMC/nmc.py:257-318 on the adjacency of self.J."""
        return nmc_core.find_clusters(self._problem_for(self.J, self.h), magnetizations, threshold_initial,
                                      threshold_cutoff, threshold_step)

    # -- dense <-> edge layout of the messages ---------------------------------------------------------
    @staticmethod
    def _lbp_pattern(J, h_msgs, u_msgs):
        """Entries the device must track explicitly: J != 0, u_msgs != 0, and every h_msgs entry that differs from
        its row's common off-pattern value (the reference's own outputs have none of those)."""
        Jd = J.toarray() if sp.issparse(J) else np.asarray(J, dtype=np.float64)
        n = Jd.shape[0]
        eye = np.eye(n, dtype=bool)
        P = (Jd != 0) | (u_msgs != 0)
        P |= P.T
        P |= eye & (h_msgs != 0)  # the diagonal of h_msgs is treated as 0 unless stored
        for _ in range(2):  # symmetrising can only add entries, so two passes settle it
            off = ~P & ~eye
            rep = h_msgs[np.arange(n), np.argmax(off, axis=1)]
            P |= off & (h_msgs != rep[:, None])
            P |= P.T
        off = ~P & ~eye
        rep = np.where(off.any(axis=1), h_msgs[np.arange(n), np.argmax(off, axis=1)], 0.0)
        return Jd, P, rep

    def LoopyBeliefPropagation(self, J, h, beta, h_msgs, u_msgs, tolerance, max_iterations):
        """One LBP call (NMC/nmc.py:168-228), kernel K5.  Returns (magnetizations, correlations, h_tilde, J_tilde,
        iteration, h_msgs, u_msgs) with the reference's dense shapes."""
        if max_iterations < 1:
            raise UnboundLocalError("cannot access local variable 'iteration' where it is not associated with a value")
        h = np.asarray(h, dtype=np.float64).reshape(-1)
        h_in = np.asarray(h_msgs, dtype=np.float64)
        u_in = np.asarray(u_msgs, dtype=np.float64)
        Jd, P, rep = self._lbp_pattern(J, h_in, u_in)
        n = Jd.shape[0]
        rows, cols = np.nonzero(P)
        A = sp.csr_matrix((Jd[rows, cols], (rows, cols)), shape=(n, n))  # keeps explicit zeros
        prob = self._problem_for_pattern(A, h)
        lbp = _lib.Lbp(prob.lbp_instance())
        try:
            r_of, c_of = prob.row_of, prob.ci
            lbp.set_messages(h_in[r_of, c_of], u_in[r_of, c_of], rep)
            marg, iteration = lbp.run(h, beta, tolerance, max_iterations)
            corr, h_tilde, J_tilde = lbp.byproducts(beta)
            he, ue, tot = lbp.get_messages()
        finally:
            lbp.close()
        H = np.repeat(tot[:, None], n, axis=1)
        np.fill_diagonal(H, 0.0)
        H[r_of, c_of] = he
        U = np.zeros((n, n))
        U[r_of, c_of] = ue
        return marg, corr, h_tilde, J_tilde, iteration, H, U

    def _problem_for_pattern(self, A: sp.csr_matrix, h) -> host.Problem:
        """Problem built from a CSR that may hold explicit zeros (host.Problem would keep them too, since it only
        re-wraps a csr_matrix); cached like any other instance."""
        A.sort_indices()
        return self._problem_for(A, h)

    def LBP_convexified(self, lambda_start, lambda_end, lambda_reduction_factor, m_star, epsilon, tolerance,
                        max_iterations, threshold_initial, threshold_cutoff, global_beta):
        """lambda-annealed LBP (NMC/nmc.py:93-166).  Returns (clusters, marginals_all_lambdas,
        mean_marginals_all_lambdas, h_tilde_all_lambdas, J_tilde_all_lambdas), dicts keyed by lambda."""
        h = np.asarray(self.h, dtype=np.float64).copy().reshape(-1)
        m_star = np.asarray(m_star, dtype=np.float64).copy().reshape(-1)
        epsilon = np.asarray(epsilon, dtype=np.float64).reshape(-1)
        prob = self._problem_for(self.J, self.h)
        marginals, means, h_tildes, J_tildes = (defaultdict(list) for _ in range(4))
        lbp = _lib.Lbp(prob.lbp_instance())
        try:
            lbp.reset(m_star)  # h_msgs = 0, u_msgs = J * m_star (nmc.py:128-129)
            lambda_val = lambda_start
            marginal = marginal_prev = None
            while lambda_val >= lambda_end:
                h_lambda = h + lambda_val * m_star * epsilon  # soft clamping at m_star (nmc.py:133-134)
                marg, iteration = lbp.run(h_lambda, global_beta, tolerance, max_iterations)
                _, h_tilde, J_tilde = lbp.byproducts(global_beta, want_corr=False)
                if iteration == max_iterations - 1 and lambda_val == lambda_start:
                    raise ValueError('LBP diverged at initial lambda, please try a larger lambda_start or increase '
                                     'max_iterations or beta')
                elif iteration == max_iterations - 1:
                    lambda_end = lambda_val
                    marginal = marginal_prev
                else:
                    marginal = marginal_prev = marg
                marginals[lambda_val] = marginal
                means[lambda_val] = np.mean(marginal)
                h_tildes[lambda_val] = h_tilde
                J_tildes[lambda_val] = J_tilde
                lambda_val = lambda_val * lambda_reduction_factor
                if round(lambda_val, 6) == 0:
                    break
        finally:
            lbp.close()
        if marginal is None:
            raise UnboundLocalError("cannot access local variable 'marginal' where it is not associated with a value")
        clusters = nmc_core.find_clusters(prob, marginal, threshold_initial, threshold_cutoff, 0.01)
        if self.verbose:
            print(f"\ncluster size = {sum(len(cluster) for cluster in clusters)}\n")
        return clusters, marginals, means, h_tildes, J_tildes

    def NMC_subroutine(self, m_star, num_cycles, num_sweeps_per_NMC_phase, full_update_frequency, M_skip, global_beta,
                       temp_x, lambda_start, lambda_end, lambda_reduction_factor, threshold_initial, threshold_cutoff,
                       max_iterations, tolerance, all_clusters=None, hash_table=None, use_hash_table=False):
        """The NMC cycle (NMC/nmc.py:320-440 for NMC, NPT/npt.py:357-477 for NPT).  Returns (M_overall,
        energy_overall, min_energy, all_clusters)."""
        prob = self._problem_for(self.J, self.h)
        N = prob.n
        m_star = np.asarray(m_star, dtype=np.float64).reshape(-1)
        _check_hash_table(hash_table, use_hash_table, num_cycles > 0 and num_sweeps_per_NMC_phase > 0 and N > 0)
        kw = dict(num_cycles=num_cycles, full_update_frequency=full_update_frequency, M_skip=M_skip,
                  global_beta=global_beta, temp_x=temp_x, lambda_start=lambda_start, lambda_end=lambda_end,
                  lambda_reduction_factor=lambda_reduction_factor, threshold_initial=threshold_initial,
                  threshold_cutoff=threshold_cutoff, max_iterations=max_iterations, tolerance=tolerance)
        if self.mode != "replay":
            from .production import nmc_subroutine_production
            return nmc_subroutine_production(self, prob, m_star, num_sweeps_per_NMC_phase, kw, self._nmc_variant,
                                             all_clusters)
        S = nmc_core.nmc_phase_count(num_cycles, full_update_frequency) * num_sweeps_per_NMC_phase
        perm, u = host.draw_sweeps(np.random, S, N)  # LBP draws nothing, so the phases' draws can be taken up front
        reps = _lib.Replicas(prob.inst, 1)
        try:
            res = nmc_core.nmc_subroutine_replay(prob, reps, m_star[None, :], variant=self._nmc_variant,
                                                 perm=perm[None], u=u[None], phase_sweeps=num_sweeps_per_NMC_phase,
                                                 all_clusters=all_clusters, verbose=self.verbose, **kw)
        finally:
            reps.close()
        return res[0]
