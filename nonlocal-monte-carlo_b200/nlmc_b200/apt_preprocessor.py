"""Drop-in replacement of the reference's ``APT_preprocessor`` (NPT/apt_preprocessor.py): adaptive
inverse-temperature ladder from the energy fluctuations sigma_E of ``num_rng`` independent chains.
Same constructor, ``run()`` keywords, return value ``(beta: list, sigma: list)`` and ``.npy`` files;
all chains of one beta run concurrently on the GPU (kernel K1 with fused per-sweep energies).
"""
from __future__ import annotations

import os
from copy import deepcopy

import numpy as np

from . import _lib, host
from .path_methods import SweepMethodsFixedInstance


class APT_preprocessor(SweepMethodsFixedInstance):
    """Reference: NPT/apt_preprocessor.py:12-31."""

    def __init__(self, J, h, *, mode: str = "replay", device: int = 0, verbose: bool = False):
        self.J = J
        if isinstance(h, list):
            h = np.array(h)
        if len(h.shape) == 1:
            h = h[:, np.newaxis]
        self.h = h
        self.N = J.shape[0]
        if mode not in ("replay", "production"):
            raise ValueError("mode must be 'replay' or 'production'")
        self.mode = mode
        self.device = device
        self.verbose = verbose

    def MCMC_task(self, m_start, beta, num_sweeps_MCMC, num_sweeps_read, use_hash_table=0):
        """One chain of the preprocessor (NPT/apt_preprocessor.py:76-113): (Energy of the last num_sweeps_read
        sweeps, final state as a (1, N) row).  The table of the reference is a memoisation and is not needed."""
        M = self.MCMC(num_sweeps_MCMC, np.asarray(m_start).copy(), beta)
        mm = M[:, -num_sweeps_read:]
        if mm.shape[1] != num_sweeps_read:  # the reference indexes column kk of a narrower matrix
            raise IndexError(f"index {mm.shape[1]} is out of bounds for axis 1 with size {mm.shape[1]}")
        prob = self._problem_for(self.J, self.h)
        Energy = prob.inst.energy_states(np.ascontiguousarray(mm.T, dtype=np.int8))
        return Energy, mm[:, -1].copy().reshape(1, -1)

    def run(self, num_sweeps_MCMC=1000, num_sweeps_read=1000, num_rng=100,
            beta_start=0.5, alpha=1.25, sigma_E_val=1000, beta_max=30, use_hash_table=1, num_cores=8):
        """APT_preprocessor.run (NPT/apt_preprocessor.py:115-204).  ``use_hash_table`` / ``num_cores`` are
        accepted and ignored.  Writes Results/data/{Energy,sigma}_iter_i.npy, beta_list_python.npy and
        sigma_list_python.npy exactly like the reference."""
        foldername = 'data'
        os.makedirs(os.path.join('Results', foldername), exist_ok=True)

        norm_factor = host.max_abs(self.J)  # apt_preprocessor.py:135-140
        self.J = self.J / norm_factor
        self.h = self.h / norm_factor
        if self.h.shape[0] == 1:
            self.h = self.h.T
        if num_sweeps_MCMC < 0 or num_sweeps_read < 0:
            # np.zeros((N, num_sweeps)) inside the worker (apt_preprocessor.py:50); pinned by the
            # reference's test_valid_parameters (NPT/unittests/test_apt_preprocessor.py:45-50)
            raise ValueError("negative dimensions are not allowed")

        prob = host.Problem(self.J, self.h, self.device)
        n = prob.n
        beta = [deepcopy(beta_start)]
        iter = 1
        sigma_E = deepcopy(sigma_E_val)
        sigma_E_min = 0.5 * np.min(np.abs(prob.val[prob.val != 0]))
        sigma = []
        saved_state = np.zeros((num_rng, n))
        reps = _lib.Replicas(prob.inst, num_rng)
        if self.mode != "replay":
            from .production import apt_preprocessor_chains_production as run_chains
        else:
            run_chains = self._chains_replay

        while sigma_E > sigma_E_min:
            if iter != 1:
                beta.append(beta[-1] + alpha / sigma_E)
            Energy, saved_state = run_chains(prob, reps, iter, saved_state, beta[-1], num_sweeps_MCMC,
                                             num_sweeps_read, num_rng)
            sigma_E = np.mean(np.std(Energy, axis=1))
            if self.verbose:
                print(f'\ncurrent iteration = {iter}, β = {beta[-1]:.3f}, and average σ = {sigma_E:.3f}\n')
            if beta[-1] > beta_max:
                if self.verbose:
                    print('Did not converge but hit the max beta limit\n')
                break
            sigma.append(sigma_E)
            np.save(os.path.join('Results', foldername, f'Energy_iter_{iter}.npy'), Energy)
            np.save(os.path.join('Results', foldername, f'sigma_iter_{iter}.npy'), sigma_E)
            iter += 1
        reps.close()
        np.save('beta_list_python.npy', beta)
        np.save('sigma_list_python.npy', sigma)
        self.plot_results(beta, sigma)
        return beta, sigma

    @staticmethod
    def _chains_replay(prob, reps, iter, saved_state, beta, num_sweeps_MCMC, num_sweeps_read, num_rng):
        """One beta iteration in exact-replay mode: the reference forks a fresh pool per iteration
        (apt_preprocessor.py:160) at the first submit, i.e. after the first chain's m_start draw."""
        n = prob.n
        m_start = np.empty((num_rng, n))
        worker = None
        for j in range(num_rng):
            if iter == 1:
                m_start[j] = np.sign(2. * np.random.rand(n, 1) - 1).reshape(-1)  # apt_preprocessor.py:164
            else:
                m_start[j] = saved_state[j]
            if worker is None:
                worker = host.fork_rng()
        sched = np.full((num_rng, num_sweeps_MCMC), float(beta))
        record_from = max(0, num_sweeps_MCMC - 1)
        Mi8, E = host.replay_chains(prob, reps, m_start, sched, worker, record_from=record_from)
        Energy = E[:, num_sweeps_MCMC - num_sweeps_read:] if num_sweeps_read else E[:, :0]  # last num_sweeps_read sweeps
        last = Mi8[:, -1, :].astype(np.float64) if num_sweeps_MCMC else m_start
        return np.ascontiguousarray(Energy), last

    def plot_results(self, beta, sigma):
        """beta_sigma.png (apt_preprocessor.py:206-231); written only when matplotlib is importable."""
        try:
            import matplotlib
            matplotlib.use("Agg")
            import matplotlib.pyplot as plt
        except Exception:
            return
        fig, ax1 = plt.subplots()
        ax1.plot(beta, marker='*', label='beta')
        ax1.set_ylabel('beta')
        ax2 = ax1.twinx()
        ax2.plot(sigma, marker='>', color='tab:orange', label='sigma')
        ax2.set_ylabel('sigma')
        ax1.set_xlabel('iteration')
        fig.savefig('beta_sigma.png')
        plt.close(fig)
