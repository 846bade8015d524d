"""Drop-in replacement of the reference's ``APT_ICM`` (NPT/apt_ICM.py): adaptive parallel tempering
with Houdayer iso-cluster moves, 10 sub-replicas per temperature.  Same constructor, ``run()``
keywords, return value ``(M, Energy)`` and attributes.  All replica x sub-replica chains of a round
run in one launch of kernel K1; the disagreement clusters of all pairs of a round come from one
launch of the connected-components kernel K7.
"""
from __future__ import annotations

import numpy as np

from . import _lib, host
from .path_methods import EnergyMethods, SweepMethodsFixedInstance


class APT_ICM(SweepMethodsFixedInstance, EnergyMethods):
    """Reference: NPT/apt_ICM.py:14-34."""

    def __init__(self, J, h, *, mode: str = "replay", device: int = 0, verbose: bool = False):
        self.J = J
        if isinstance(h, list):
            h = np.array(h)
        if len(h.shape) == 1:
            h = h[:, np.newaxis]
        self.h = h
        if mode not in ("replay", "production"):
            raise ValueError("mode must be 'replay' or 'production'")
        self.mode = mode
        self.device = device
        self.verbose = verbose

    def select_non_overlapping_pairs(self, all_pairs):
        return host.select_non_overlapping_pairs(all_pairs, self.num_swapping_pairs)

    def find_disagreement_clusters(self, state_1, state_2, J=None):
        """apt_ICM.py:116-143 through kernel K7; returns the reference's list of lists (clusters ordered
        by smallest site index; members listed in increasing order)."""
        prob = self._problem_for(self.J if J is None else J, self.h)
        labels, counts = _lib.icm_clusters(prob.inst, host.as_spins_i8(state_1), host.as_spins_i8(state_2))
        return [list(np.flatnonzero(labels[0] == k)) for k in range(int(counts[0]))]

    def run(self, beta_list, num_replicas, num_sweeps_MCMC=1000, num_sweeps_read=1000, num_swap_attempts=100,
            num_swapping_pairs=1, use_hash_table=0, num_cores=8):
        """APT_ICM.run (NPT/apt_ICM.py:145-305).  J and h are used as given (the reference does not
        normalise inside run; its caller does, apt_ICM.py:342-344)."""
        self.num_replicas = num_replicas
        self.num_sweeps_MCMC = num_sweeps_MCMC
        self.num_sweeps_read = num_sweeps_read
        self.num_swap_attempts = num_swap_attempts
        self.num_sweeps_MCMC_per_swap = self.num_sweeps_MCMC // self.num_swap_attempts
        self.num_sweeps_read_per_swap = self.num_sweeps_read // self.num_swap_attempts
        self.num_swapping_pairs = num_swapping_pairs
        self.use_hash_table = use_hash_table
        if self.mode != "replay":
            from .production import apt_icm_run_production
            return apt_icm_run_production(self, np.asarray(beta_list, dtype=np.float64))

        num_subreplicas = 10  # apt_ICM.py:177
        useKatzgraber = True
        S, R, spm, spr = num_subreplicas, num_replicas, self.num_sweeps_MCMC_per_swap, self.num_sweeps_read_per_swap
        if spm < 0:
            raise ValueError("negative dimensions are not allowed")
        beta_list = np.asarray(beta_list, dtype=np.float64)
        prob = host.Problem(self.J, self.h, self.device)
        n = prob.n
        count = np.zeros(self.num_swap_attempts)
        all_pairs = [(i, i + 1) for i in range(1, R)]
        reps = _lib.Replicas(prob.inst, R * S)
        M = np.zeros((n * R, spm * S))
        E_cols = np.zeros((R, S, spm))
        m_start_matrix = np.sign(2 * np.random.rand(n * R, S) - 1)  # apt_ICM.py:188
        sched = np.repeat(np.repeat(beta_list[:R], S)[:, None], spm, axis=1)  # chain g = replica*S + sub

        for ii in range(int(self.num_swap_attempts)):
            if self.verbose:
                print(f"\nRunning swap attempt = {ii + 1}")
            # MCMC for every (replica, sub-replica), consuming np.random in that order (apt_ICM.py:197-213)
            starts = m_start_matrix.reshape(R, n, S).transpose(0, 2, 1).reshape(R * S, n)
            Mi8, E = host.replay_chains(prob, reps, starts, sched, np.random)
            for r in range(R):
                for s in range(S):
                    g = r * S + s
                    M[r * n:(r + 1) * n, s * spm:(s + 1) * spm] = Mi8[g].T
                    m_start_matrix[r * n:(r + 1) * n, s] = Mi8[g][-1] if spm else m_start_matrix[r * n:(r + 1) * n, s]
            E_cols[:] = E.reshape(R, S, spm)

            # Houdayer move on the FIRST column of each block (apt_ICM.py:216-246).  np.random.randint
            # consumes a data-dependent amount of the stream, so replica r+1's pairing depends on replica
            # r's cluster counts: the pairs are processed replica by replica.
            for r in range(R):
                shuffled = np.random.permutation(S)
                rows = slice(r * n, (r + 1) * n)
                pairs = [(int(shuffled[2 * p]), int(shuffled[2 * p + 1])) for p in range(S // 2)]
                # the five pairs of a replica touch disjoint sub-replicas: one K7 launch for all of them
                s1 = np.stack([M[rows, a * spm] for a, _ in pairs])
                s2 = np.stack([M[rows, b * spm] for _, b in pairs])
                labels, counts = _lib.icm_clusters(prob.inst, s1.astype(np.int8), s2.astype(np.int8))
                edited = []
                for p, (a, b) in enumerate(pairs):
                    if not counts[p]:
                        continue
                    pick = np.random.randint(int(counts[p]))  # apt_ICM.py:233
                    members = labels[p] == pick
                    state_1, state_2 = s1[p].copy(), s2[p].copy()
                    if useKatzgraber and int(members.sum()) > n // 2:  # apt_ICM.py:236-237
                        state_1 = -state_1
                    else:
                        state_1[members], state_2[members] = s2[p][members], s1[p][members]
                    M[rows, a * spm] = state_1
                    M[rows, b * spm] = state_2
                    edited += [(a, state_1), (b, state_2)]
                if edited:  # energies of the edited first columns (they can enter the returned Energy)
                    E_new = prob.inst.energy_states(np.stack([st for _, st in edited]).astype(np.int8))
                    for (sub, _), e in zip(edited, E_new):
                        E_cols[r, sub, 0] = e

            selected_pairs = self.select_non_overlapping_pairs(all_pairs)
            for s in range(S):  # PT swap per sub-replica on the LAST column of its block (apt_ICM.py:251-285)
                for sel, nxt in selected_pairs:
                    E_sel, E_next = E_cols[sel - 1, s, spm - 1], E_cols[nxt - 1, s, spm - 1]
                    DeltaE = E_next - E_sel
                    DeltaB = beta_list[nxt - 1] - beta_list[sel - 1]
                    if np.random.rand() < min(1, np.exp(DeltaB * DeltaE)):
                        count[ii] += 1
                        col = (s + 1) * spm - 1
                        m_start_matrix[(sel - 1) * n:sel * n, s] = M[(nxt - 1) * n:nxt * n, col]
                        m_start_matrix[(nxt - 1) * n:nxt * n, s] = M[(sel - 1) * n:sel * n, col]
        reps.close()

        # apt_ICM.py:291-297: minimum over the first num_sweeps_read_per_swap columns of M
        Energy = np.zeros(R)
        E_flat = E_cols.reshape(R, S * spm)
        self._EE1_list = []
        for r in range(R):
            EE1 = E_flat[r, :spr].copy()
            Energy[r] = np.min(EE1)
            self._EE1_list.append(EE1)
        if self.verbose:
            print(f"\nLatest energy from each replica = {Energy}")
            print(f"Swap acceptance rate = {np.count_nonzero(count) / max(count.size, 1) * 100:.2f} per cent\n")
        self.plot_energies(self._EE1_list, beta_list)
        return M, Energy

    def plot_energies(self, EE1_list, beta_list):
        """'APT_ICM_energy..png' (NPT/apt_ICM.py:307-322, the reference's file name); written only when matplotlib
        is importable."""
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
        plt.savefig('APT_ICM_energy..png')
        plt.close()
