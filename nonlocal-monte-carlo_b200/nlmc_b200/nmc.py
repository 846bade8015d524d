"""Drop-in replacement of the reference's ``NMC`` class (NMC/nmc.py): Non-equilibrium (nonlocal)
Monte Carlo.  Same constructor, same positional/keyword arguments of ``run()``, same return tuple
``(M_overall, energy_overall, min_energy)``; sweeps, energies and LBP run in libnlmc_b200.so.
"""
from __future__ import annotations

import numpy as np

from . import _lib, host
from .nmc_core import nmc_phase_count, nmc_subroutine_replay
from .path_methods import LbpMethods, SweepMethods


class NMC(SweepMethods, LbpMethods):
    """Reference: NMC/nmc.py:13-26.  MCMC, LBP_convexified, LoopyBeliefPropagation, atanh_saturated, find_clusters and
    NMC_subroutine come from the mixins in path_methods.py."""

    _nmc_variant = "nmc"

    def __init__(self, J, h, *, mode: str = "replay", device: int = 0, verbose: bool = False):
        self.J = J
        self.h = h
        self.h = np.asarray(h).reshape(-1)
        if mode not in ("replay", "production"):
            raise ValueError("mode must be 'replay' or 'production'")
        self.mode = mode
        self.device = device
        self.verbose = verbose

    def run(self, num_sweeps_initial=int(1e4), num_sweeps_per_NMC_phase=int(1e4),
            num_NMC_cycles=10, full_update_frequency=1, M_skip=1, temp_x=20,
            global_beta=2.5, lambda_start=0.5, lambda_end=0.01, lambda_reduction_factor=0.9,
            threshold_initial=0.999999, threshold_cutoff=0.99999, max_iterations=100, tolerance=np.finfo(float).eps,
            use_hash_table=False):
        """NMC.run (NMC/nmc.py:442-520).  ``use_hash_table`` is accepted and ignored (a CPU memoisation
        with no effect on results).  Returns (M_overall, energy_overall, min_energy)."""
        norm_factor = host.max_abs(self.J)  # nmc.py:472-476
        self.J = self.J / norm_factor
        self.h = self.h / norm_factor
        if self.mode != "replay":
            from .production import nmc_run_production
            return nmc_run_production(self, dict(
                num_sweeps_initial=num_sweeps_initial, num_sweeps_per_NMC_phase=num_sweeps_per_NMC_phase,
                num_NMC_cycles=num_NMC_cycles, full_update_frequency=full_update_frequency, M_skip=M_skip,
                temp_x=temp_x, global_beta=global_beta, lambda_start=lambda_start, lambda_end=lambda_end,
                lambda_reduction_factor=lambda_reduction_factor, threshold_initial=threshold_initial,
                threshold_cutoff=threshold_cutoff, max_iterations=max_iterations, tolerance=tolerance))
        N = len(self.h)
        if num_sweeps_initial < 0 or num_sweeps_per_NMC_phase < 0:
            raise ValueError("negative dimensions are not allowed")
        prob = host.Problem(self.J, self.h, self.device)
        reps = _lib.Replicas(prob.inst, 1)

        m_init = np.sign(2 * np.random.rand(N) - 1)  # nmc.py:487
        # annealed MCMC from beta 0 to global_beta to find m_star (nmc.py:490-502)
        sched = host.beta_schedule(num_sweeps_initial, global_beta, anneal=True, sweeps_per_beta=1, initial_beta=0)
        Mi8, E = host.replay_chains(prob, reps, m_init[None, :], sched[None, :], np.random)
        Energy_star = np.min(E[0])
        m_star = Mi8[0][int(np.argmin(E[0]))].astype(np.float64)
        if self.verbose:
            print(f'\ninitial m_star energy = {Energy_star:.8f}')

        # the phases' draws do not depend on the spins: each phase draws its own sweeps (same stream order as the
        # reference, one phase of permutations and uniforms in host memory at a time)
        res = nmc_subroutine_replay(prob, reps, m_star[None, :], variant="nmc", perm=None, u=None,
                                    draw=lambda n_sweeps: host.draw_sweeps(np.random, n_sweeps, N),
                                    num_cycles=num_NMC_cycles, phase_sweeps=num_sweeps_per_NMC_phase,
                                    full_update_frequency=full_update_frequency, M_skip=M_skip,
                                    global_beta=global_beta, temp_x=temp_x, lambda_start=lambda_start,
                                    lambda_end=lambda_end, lambda_reduction_factor=lambda_reduction_factor,
                                    threshold_initial=threshold_initial, threshold_cutoff=threshold_cutoff,
                                    max_iterations=max_iterations, tolerance=tolerance, verbose=self.verbose)
        M_overall, energy_overall, min_energy, all_clusters = res[0]
        reps.close()
        self.all_clusters = all_clusters
        self.plot_results(M_overall, energy_overall, all_clusters, M_skip, num_NMC_cycles, full_update_frequency,
                          num_sweeps_per_NMC_phase)
        return M_overall, energy_overall, min_energy

    def plot_results(self, M_overall, energy_overall, all_clusters, M_skip, num_NMC_cycles, full_update_frequency,
                     num_sweeps_per_NMC_phase):
        """NMC_spins.png / NMC_energy.png (NMC/nmc.py:522-641); written only when matplotlib is importable."""
        try:
            import matplotlib
            matplotlib.use("Agg")
            import matplotlib.pyplot as plt
        except Exception:
            return
        fig, axes = plt.subplots(2, 1, figsize=(10, 6))
        axes[0].imshow(M_overall[all_clusters, ::M_skip], aspect='auto', cmap='viridis')
        axes[0].set_title('backbone spins')
        rest = np.setdiff1d(np.arange(M_overall.shape[0]), all_clusters)
        axes[1].imshow(M_overall[rest, ::M_skip], aspect='auto', cmap='viridis')
        axes[1].set_title('non-backbone spins')
        fig.savefig('NMC_spins.png')
        plt.close(fig)
        fig = plt.figure()
        plt.plot(energy_overall)
        plt.xlabel('stored sweep')
        plt.ylabel('Energy')
        fig.savefig('NMC_energy.png')
        plt.close(fig)
