"""Host driver of the NMC building blocks shared by ``NMC.run`` and the ``doNMC`` replicas of
``NPT.run``: the lambda schedule around the LBP kernel (K5), backbone selection, and the three-phase
NMC cycle executed with kernel K1 on a batch of chains.

What runs where: every sweep, every energy and every LBP iteration is a CUDA kernel.  The host
keeps what the reference does in a handful of scalar operations per call: the lambda loop and its
divergence rules (NMC/nmc.py:131-161), ``find_clusters`` (set operations on the seed spins,
nmc.py:257-318), and the per-phase bookkeeping ``m_init = M[:, argmin E]`` (nmc.py:394-395).
"""
from __future__ import annotations

import numpy as np

from . import _lib, host


def nmc_phase_count(num_cycles: int, full_update_frequency: int) -> int:
    """Number of MCMC phases one NMC_subroutine call executes (NMC/nmc.py:365-421)."""
    return sum(2 + (1 if cycle % full_update_frequency == 0 else 0) for cycle in range(num_cycles))


def _adjacency(prob):
    """Per site the sorted distinct neighbours with a non-zero coupling (np.where(J[i, :] != 0)[0]), cached."""
    adj = getattr(prob, "_adjacency_lists", None)
    if adj is None:
        adj = []
        for i in range(prob.n):
            b, e = prob.rp[i], prob.rp[i + 1]
            adj.append(sorted(set(prob.ci[b:e][prob.val[b:e] != 0].tolist())))
        try:
            prob._adjacency_lists = adj
        except AttributeError:
            pass
    return adj


def find_clusters(prob: host.Problem, magnetizations, threshold_initial, threshold_cutoff, threshold_step):
    """Backbone seeds and growth with the semantics of find_clusters (NMC/nmc.py:257-318): seeds are the
    spins with |marginal| >= threshold_initial; a seed not yet clustered opens a cluster with its
    unclustered seed neighbours (in increasing index order); clusters then grow over unclustered neighbours
    whose |marginal| is above a threshold lowered by threshold_step until it reaches threshold_cutoff."""
    absm = np.abs(np.asarray(magnetizations, dtype=np.float64))
    adj = _adjacency(prob)
    is_seed = (absm >= threshold_initial).tolist()
    clustered = [False] * prob.n
    clusters = []
    for seed in np.flatnonzero(absm >= threshold_initial).tolist():
        if clustered[seed]:
            continue
        members = [seed] + [j for j in adj[seed] if is_seed[j] and not clustered[j]]
        for j in members:
            clustered[j] = True
        clusters.append(members)
    current = threshold_initial - threshold_step
    while current > threshold_cutoff:
        for i, members in enumerate(clusters):
            cand = sorted({j for k in members for j in adj[k] if not clustered[j]})
            grown = [j for j in cand if absm[j] >= current]
            for j in grown:
                clustered[j] = True
            clusters[i] = members + grown
        current -= threshold_step
    return [np.array(c, dtype=int) for c in clusters]


def backbone_indices(prob: host.Problem, marginal, threshold_initial, threshold_cutoff, threshold_step=0.01):
    """The backbone as a sorted index set: what NMC_subroutine uses of find_clusters' result (J_c[all_clusters, :],
    NMC/nmc.py:373-381).  With every shipped parameter set the growth loop never runs (threshold_initial - step <=
    threshold_cutoff) and the set is simply {i : |marginal_i| >= threshold_initial}."""
    if threshold_initial - threshold_step <= threshold_cutoff:
        return np.flatnonzero(np.abs(np.asarray(marginal, dtype=np.float64)) >= threshold_initial).astype(int)
    cl = find_clusters(prob, marginal, threshold_initial, threshold_cutoff, threshold_step)
    return np.sort(np.concatenate(cl)).astype(int) if cl else np.array([], dtype=int)


def lbp_convexified(prob: host.Problem, lbp: "_lib.Lbp", m_star, lambda_start, lambda_end, lambda_reduction_factor,
                    tolerance, max_iterations, threshold_initial, threshold_cutoff, global_beta, trace=None,
                    as_index_set=False):
    """lambda-annealed LBP (NMC/nmc.py:93-166) around kernel K5.  Returns the list of clusters (or, with
    as_index_set, the backbone as a sorted index array -- see backbone_indices).
    ``trace`` (optional list) receives (lambda, iteration, marginal) per step, for tests."""
    lbp.reset(m_star)
    lambda_val = lambda_start
    marginal = None
    marginal_prev = None
    while lambda_val >= lambda_end:
        marg, iteration = lbp.step(lambda_val, global_beta, tolerance, max_iterations)
        if trace is not None:
            trace.append((lambda_val, iteration, marg.copy()))
        if iteration == max_iterations - 1 and lambda_val == lambda_start:
            raise ValueError(
                'LBP diverged at initial lambda, please try a larger lambda_start or increase max_iterations or beta')
        elif iteration == max_iterations - 1:
            lambda_end = lambda_val
            marginal = marginal_prev
        else:
            marginal = marg
            marginal_prev = marg
        lambda_val = lambda_val * lambda_reduction_factor
        if round(lambda_val, 6) == 0:
            break
    if marginal is None:  # lambda_start < lambda_end: the reference fails on an unbound `marginal`
        raise UnboundLocalError("cannot access local variable 'marginal' where it is not associated with a value")
    if as_index_set:
        return backbone_indices(prob, marginal, threshold_initial, threshold_cutoff, 0.01)
    return find_clusters(prob, marginal, threshold_initial, threshold_cutoff, 0.01)


def nmc_subroutine_replay(prob: host.Problem, reps: "_lib.Replicas", m_star, *, variant: str, perm, u,
                          num_cycles, phase_sweeps, full_update_frequency, M_skip, global_beta, temp_x,
                          lambda_start, lambda_end, lambda_reduction_factor, threshold_initial, threshold_cutoff,
                          max_iterations, tolerance, all_clusters=None, verbose=False, draw=None):
    """NMC_subroutine for a batch of G chains in lock step (exact-replay mode).

    variant "nmc": NMC/nmc.py:320-440 (LBP inside the cycle loop, m_star follows the ALL phase);
    variant "npt": NPT/npt.py:357-477 (LBP once, before the loop).
    perm/u [G][n_phases*phase_sweeps][n]: the reference's draws for chain g, in phase order; or None with
    draw(n_sweeps) -> (perm, u) [n_sweeps][n], called once per phase (single chain only: the draws do not depend
    on the spins, so drawing a phase at a time consumes the stream in the same order with one phase in memory).
    Returns a list of (M_overall float64 [n][cols], energy_overall, min_energy, all_clusters) per chain.
    """
    G, n = reps.R, prob.n
    m_init = np.asarray(m_star, dtype=np.float64).reshape(G, n).copy()
    m_star = m_init.copy()
    cap = phase_sweeps * num_cycles * 3 // M_skip
    M_overall = [np.zeros((n, cap)) for _ in range(G)]
    E_overall = [np.zeros(cap) for _ in range(G)]
    M_index = 0
    width = phase_sweeps // M_skip
    beta_sched = np.full((G, phase_sweeps), float(global_beta))
    lut = prob.tanh_lut(beta_sched)
    lbp = None
    clusters_provided = all_clusters is not None
    clusters = [np.asarray(all_clusters, dtype=int)] * G if clusters_provided else [None] * G
    phase_no = 0

    def backbone(g):
        nonlocal lbp
        if lbp is None:
            lbp = _lib.Lbp(prob.lbp_instance())
        cl = lbp_convexified(prob, lbp, m_star[g], lambda_start, lambda_end, lambda_reduction_factor, tolerance,
                             max_iterations, threshold_initial, threshold_cutoff, global_beta)
        if verbose:
            print(f"\ncluster size = {sum(len(c) for c in cl)}\n")
        return np.concatenate(cl).astype(int) if cl else np.array([], dtype=int)

    def run_phase():
        nonlocal phase_no, M_index, m_init
        reps.set_spins(m_init)
        if perm is None:
            assert G == 1 and draw is not None
            p1, u1 = draw(phase_sweeps)
            Mi8, E = reps.sweep_replay(p1[None], u1[None], beta_sched, lut, prob.lut_half)
        else:
            sl = slice(phase_no * phase_sweeps, (phase_no + 1) * phase_sweeps)
            Mi8, E = reps.sweep_replay(perm[:, sl], u[:, sl], beta_sched, lut, prob.lut_half)
        phase_no += 1
        for g in range(G):
            Mg = Mi8[g].T.astype(np.float64)
            M_overall[g][:, M_index:M_index + width] = Mg[:, ::M_skip]
            E_overall[g][M_index:M_index + width] = E[g][::M_skip]
            m_init[g] = Mg[:, int(np.argmin(E[g]))]  # first minimum wins (np.argmin, nmc.py:394-395)
        M_index += width
        return E

    if variant == "npt" and not clusters_provided:
        clusters = [backbone(g) for g in range(G)]
    for cycle in range(num_cycles):
        if verbose:
            print(f'\nCurrent iteration = {cycle + 1}')
        if variant == "nmc" and not clusters_provided:
            clusters = [backbone(g) for g in range(G)]
        in_cl = np.zeros((G, n), dtype=bool)
        for g in range(G):
            in_cl[g, clusters[g]] = True
        # phase C: backbone rows at beta/temp_x, everything else frozen by +-1e4 (nmc.py:377-385)
        for g in range(G):
            h_c = prob.h.copy()
            h_c[in_cl[g]] /= temp_x
            h_c[~in_cl[g]] = m_init[g][~in_cl[g]] * 10000
            reps.set_phase(g, h_c, in_cl[g].astype(np.uint8), temp_x)
        run_phase()
        # phase NC: backbone frozen, the rest at beta (nmc.py:398-406)
        for g in range(G):
            h_nc = prob.h.copy()
            h_nc[in_cl[g]] = m_init[g][in_cl[g]] * 10000
            reps.set_phase(g, h_nc, None, 1.0)
        run_phase()
        # phase ALL every full_update_frequency cycles (nmc.py:419-433)
        if cycle % full_update_frequency == 0:
            for g in range(G):
                reps.set_phase(g, None, None, 1.0)
            E = run_phase()
            if variant == "nmc":
                m_star = m_init.copy()
                if verbose:
                    print(f'\ncurrent m_star energy = {np.min(E[0]):.8f}')
    for g in range(G):
        reps.set_phase(g, None, None, 1.0)
    if lbp is not None:
        lbp.close()
    out = []
    for g in range(G):
        Mo, Eo = M_overall[g][:, :M_index], E_overall[g][:M_index]
        out.append((Mo, Eo, np.min(Eo), clusters[g]))
    return out
