"""Drop-in replacement of the reference's ``NPT`` class (NPT/npt.py): Non-equilibrium Monte Carlo +
Adaptive Parallel Tempering.  Same constructor, same ``run()`` keyword arguments, same numpy
return values and attributes; the sweeps, energies, LBP backbone search and swaps run as sm_100a
CUDA kernels in libnlmc_b200.so (no CPU fallback).

Modes (constructor keyword, not part of the reference API):
  ``mode="replay"``      bit-exact reproduction of the reference run with ``num_cores=1`` for the same
                         ``np.random.seed`` / ``random.seed``: the host draws the reference's random
                         stream in the reference's order and injects it into kernel K1.
  ``mode="production"``  Philox-driven, graph-coloured, bit-packed kernels (statistically equivalent);
                         ``num_runs`` independent ladders run side by side in the bit lanes.
"""
from __future__ import annotations

import numpy as np

from . import _lib, host
from .nmc_core import nmc_phase_count, nmc_subroutine_replay
from .path_methods import EnergyMethods, LbpMethods, SweepMethods


class NPT(SweepMethods, LbpMethods, EnergyMethods):
    """Non-equilibrium Monte Carlo + Adaptive Parallel Tempering (reference: NPT/npt.py:15-29)."""

    def __init__(self, J, h, *, mode: str = "replay", device: int = 0, verbose: bool = False):
        self.J = J
        self.h = h
        self.h = np.asarray(h).reshape(-1)  # NPT/npt.py:29
        if mode not in ("replay", "production"):
            raise ValueError("mode must be 'replay' or 'production'")
        self.mode = mode
        self.device = device
        self.verbose = verbose

    # ------------------------------------------------------------------------------------------
    _nmc_variant = "npt"

    def MCMC_task(self, replica_i, num_sweeps_MCMC, m_start, beta_list, use_hash_table=False, hash_table=None):
        """NPT/npt.py:112-127: plain sweeps of replica `replica_i` (1-based) at beta_list[replica_i - 1]."""
        return self.MCMC(num_sweeps_MCMC, np.asarray(m_start).copy(), beta_list[replica_i - 1], self.J, self.h,
                         hash_table=hash_table, use_hash_table=use_hash_table)

    def NMC_task(self, m_start, num_cycles, num_sweeps_per_NMC_phase, full_update_frequency, M_skip, global_beta,
                 temp_x, lambda_start, lambda_end, lambda_reduction_factor, threshold_initial, threshold_cutoff,
                 max_iterations, tolerance, use_hash_table=False, hash_table=None):
        """NPT/npt.py:479-512: NMC_subroutine, returning only M_overall."""
        M_overall, _, _, _ = self.NMC_subroutine(
            m_start, num_cycles, num_sweeps_per_NMC_phase, full_update_frequency, M_skip, global_beta, temp_x,
            lambda_start, lambda_end, lambda_reduction_factor, threshold_initial, threshold_cutoff, max_iterations,
            tolerance, hash_table=hash_table, use_hash_table=use_hash_table)
        return M_overall

    def select_non_overlapping_pairs(self, all_pairs):
        return host.select_non_overlapping_pairs(all_pairs, self.num_swapping_pairs)

    def _problem(self):
        key = (id(self.J), id(self.h))
        if getattr(self, "_prob_key", None) != key:
            self._prob = host.Problem(self.J, self.h, self.device)
            self._prob_key = key
        return self._prob

    # ------------------------------------------------------------------------------------------
    def run(self, beta_list, num_replicas, doNMC, num_sweeps_MCMC=1000, num_sweeps_read=1000, num_swap_attempts=100,
            num_swapping_pairs=1, num_cycles=10, full_update_frequency=1, M_skip=1, temp_x=20,
            global_beta=2.5, lambda_start=0.5, lambda_end=0.01, lambda_reduction_factor=0.9,
            threshold_initial=0.999999, threshold_cutoff=0.99999, max_iterations=100, tolerance=np.finfo(float).eps,
            use_hash_table=False, num_cores=8):
        """Run NPT (reference: NPT/npt.py:535-700).  ``use_hash_table`` and ``num_cores`` are accepted and
        ignored: the hash table is a pure CPU memoisation with no effect on results and the replicas
        run concurrently on the GPU.  Returns (M, Energy) exactly as the reference does."""
        self.num_replicas = num_replicas
        self.num_sweeps_MCMC = num_sweeps_MCMC
        self.num_sweeps_read = num_sweeps_read
        self.num_swap_attempts = num_swap_attempts
        self.num_sweeps_MCMC_per_swap = self.num_sweeps_MCMC // self.num_swap_attempts
        self.num_sweeps_read_per_swap = self.num_sweeps_read // self.num_swap_attempts
        self.num_sweeps_per_NMC_phase_per_swap = int(
            np.ceil(self.num_sweeps_MCMC / self.num_swap_attempts / 3 / num_cycles))
        self.num_swapping_pairs = num_swapping_pairs
        self.use_hash_table = use_hash_table
        self.doNMC = doNMC
        self.hash_table = None

        norm_factor = host.max_abs(self.J)  # NPT/npt.py:588-590 (rebinds, never writes the caller's arrays)
        self.J, self.h = host.normalised(self.J, self.h, norm_factor)

        if len(self.doNMC) != self.num_replicas:
            raise ValueError("The length of doNMC does not match the number of replicas.")
        if self.num_sweeps_MCMC_per_swap < 0:
            raise ValueError("negative dimensions are not allowed")

        nmc_kw = dict(num_cycles=num_cycles, phase_sweeps=self.num_sweeps_per_NMC_phase_per_swap,
                      full_update_frequency=full_update_frequency, M_skip=M_skip, global_beta=global_beta,
                      temp_x=temp_x, lambda_start=lambda_start, lambda_end=lambda_end,
                      lambda_reduction_factor=lambda_reduction_factor, threshold_initial=threshold_initial,
                      threshold_cutoff=threshold_cutoff, max_iterations=max_iterations, tolerance=tolerance)
        if self.mode == "replay":
            M, Energy, count = self._run_replay(np.asarray(beta_list, dtype=np.float64), nmc_kw)
        else:
            from .production import npt_run_production
            M, Energy, count = npt_run_production(self, np.asarray(beta_list, dtype=np.float64), nmc_kw)

        if self.verbose:
            print(f"\nLatest energy from each replica = {Energy}")
            print(f"Swap acceptance rate = {np.count_nonzero(count) / max(count.size, 1) * 100:.2f} per cent\n")
        self.plot_energies(getattr(self, "_EE1_list", []), beta_list)
        return M, Energy

    # ------------------------------------------------------------------------------------------
    def _run_replay(self, beta_list, nmc_kw):
        R = self.num_replicas
        spm = self.num_sweeps_MCMC_per_swap
        spr = self.num_sweeps_read_per_swap
        prob = self._problem()
        n = prob.n
        count = np.zeros(self.num_swap_attempts)
        all_pairs = [(i, i + 1) for i in range(1, R)]

        mc_ids = [r for r in range(R) if not self.doNMC[r]]
        nmc_ids = [r for r in range(R) if self.doNMC[r]]
        mc_reps = _lib.Replicas(prob.inst, len(mc_ids)) if mc_ids else None
        nmc_reps = _lib.Replicas(prob.inst, len(nmc_ids)) if nmc_ids else None

        M = np.zeros((R * n, spm))
        E_cols = np.zeros((R, spm))  # energy of every column of M (kernel-computed)
        m_start = np.sign(2 * np.random.rand(R * n, 1) - 1)  # NPT/npt.py:612
        worker = None
        n_nmc_sweeps = nmc_phase_count(nmc_kw["num_cycles"], nmc_kw["full_update_frequency"]) * nmc_kw["phase_sweeps"]

        for ii in range(self.num_swap_attempts):
            if self.verbose:
                print(f"\nRunning swap attempt = {ii + 1}")
            if worker is None:
                worker = host.fork_rng()  # the single pool worker is forked at the first submit
            # the worker executes the R tasks in submission order, each consuming its share of the stream
            draws = {}
            for r in range(R):
                draws[r] = host.draw_sweeps(worker, n_nmc_sweeps if self.doNMC[r] else spm, n)
            ms = m_start.reshape(R, n)
            if mc_ids:
                perm = np.stack([draws[r][0] for r in mc_ids])
                u = np.stack([draws[r][1] for r in mc_ids])
                beta_sched = np.repeat(beta_list[mc_ids][:, None], spm, axis=1)
                mc_reps.set_spins(ms[mc_ids])
                Mi8, E = mc_reps.sweep_replay(perm, u, beta_sched, prob.tanh_lut(beta_sched), prob.lut_half)
                for g, r in enumerate(mc_ids):
                    M[r * n:(r + 1) * n, :] = Mi8[g].T
                    E_cols[r] = E[g]
            if nmc_ids:
                res = nmc_subroutine_replay(prob, nmc_reps, ms[nmc_ids], variant="npt",
                                            perm=np.stack([draws[r][0] for r in nmc_ids]),
                                            u=np.stack([draws[r][1] for r in nmc_ids]), **nmc_kw)
                for g, r in enumerate(nmc_ids):
                    M_overall, E_overall = res[g][0], res[g][1]
                    M[r * n:(r + 1) * n, :] = M_overall[:, -spm:]  # NPT/npt.py:643-644
                    E_cols[r] = E_overall[-spm:]

            m_start = M[:, -1].copy().reshape(-1, 1)
            selected_pairs = self.select_non_overlapping_pairs(all_pairs)
            for sel, nxt in selected_pairs:  # NPT/npt.py:652-680
                E_sel, E_next = E_cols[sel - 1, -1], E_cols[nxt - 1, -1]
                DeltaE = E_next - E_sel
                DeltaB = beta_list[nxt - 1] - beta_list[sel - 1]
                if np.random.rand() < min(1, np.exp(DeltaB * DeltaE)):
                    count[ii] += 1
                    m_sel = m_start[(sel - 1) * n:sel * n].copy()
                    m_start[(sel - 1) * n:sel * n] = m_start[(nxt - 1) * n:nxt * n]
                    m_start[(nxt - 1) * n:nxt * n] = m_sel

        # NPT/npt.py:686-692: minimum over the FIRST num_sweeps_read_per_swap columns of the last round
        Energy = np.zeros(R)
        self._EE1_list = []
        for r in range(R):
            EE1 = E_cols[r, :spr].copy()
            Energy[r] = np.min(EE1)
            self._EE1_list.append(EE1)
        for reps in (mc_reps, nmc_reps):
            if reps is not None:
                reps.close()
        return M, Energy, count

    # ------------------------------------------------------------------------------------------
    def plot_energies(self, EE1_list, beta_list):
        """NPT/npt.py:702-717; written only when matplotlib is importable."""
        try:
            import matplotlib
            matplotlib.use("Agg")
            import matplotlib.pyplot as plt
        except Exception:
            return
        plt.figure()
        for i, EE1 in enumerate(EE1_list):
            plt.plot(EE1, label=f"Replica {i + 1} (β={beta_list[i]:.2f})")
        plt.xlabel('Sweeps')
        plt.ylabel('Energy')
        plt.title('Energy traces for different replicas')
        plt.legend()
        plt.savefig('NPT_energy.png')
        plt.close()
