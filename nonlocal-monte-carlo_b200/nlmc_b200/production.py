"""Production mode of the drop-in classes: Philox-driven, graph-coloured, bit-packed kernels
(nlmc_msc_* in include/nlmc_b200.h).  Statistically equivalent to the reference (same single-site
heat-bath conditional, same swap rule), not bit-identical: the visiting order is colour-parallel and
the random stream is Philox4x32-10 instead of numpy's MT19937.

Independent runs ride in the bit lanes: a call with ``num_runs=k`` (attribute of the object, default 1)
simulates k independent ladders at once -- the lanes are padded to a multiple of 128, so up to 128
runs cost the same as one.  ``run()`` returns the reference's tuple for run 0 and keeps every run's
energies in ``self.energies_all_runs`` ([n_beta][num_runs]).
"""
from __future__ import annotations

import numpy as np

from . import _lib, host


def _seed_from_numpy() -> int:
    """A 64-bit Philox seed drawn from the global np.random, so np.random.seed(s) makes runs repeatable."""
    hi, lo = np.random.randint(0, 2**31 - 1, size=2)
    return (int(hi) << 32) | int(lo)


def _require_msc(prob: host.Problem, betas, n_ladders: int, seed: int) -> "_lib.Msc":
    try:
        return _lib.Msc(prob.inst, betas, n_ladders, seed)
    except _lib.NlmcError as e:
        raise NotImplementedError(
            "mode='production' currently covers +-J instances with h = 0 and even degrees <= 6 "
            f"(2D/3D lattices); use mode='replay' for this instance ({e})") from e


def npt_run_production(obj, beta_list, nmc_kw):
    """NPT.run (NPT/npt.py:535-700) on the bit-packed path.  Returns (M, Energy, count) for run 0."""
    if any(obj.doNMC):
        raise NotImplementedError("mode='production' does not run doNMC replicas yet; use mode='replay'")
    R = obj.num_replicas
    spm, spr = obj.num_sweeps_MCMC_per_swap, obj.num_sweeps_read_per_swap
    num_runs = int(getattr(obj, "num_runs", 1))
    prob = obj._problem()
    n = prob.n
    msc = _require_msc(prob, beta_list[:R], num_runs, _seed_from_numpy())
    count = np.zeros(obj.num_swap_attempts)
    # all rounds but the last run fused on the device; nothing comes back to the host
    for ii in range(obj.num_swap_attempts - 1):
        msc.round(spm, obj.num_swapping_pairs)
    before = msc.swap_count(reset=True) if obj.num_swap_attempts > 1 else 0
    count[:max(obj.num_swap_attempts - 1, 0)] = before / max(obj.num_swap_attempts - 1, 1)
    # last round: record the state after every sweep (the reference returns the last round's M)
    M = np.zeros((R * n, spm))
    E_cols = np.zeros((R, spm))
    E_all = None
    for j in range(spm):
        msc.sweep(1)
        E_all = msc.energies()
        E_cols[:, j] = E_all[:, 0]
        for r in range(R):
            M[r * n:(r + 1) * n, j] = msc.get_spins(r, 0)
    if obj.num_swap_attempts > 0 and spm > 0:
        msc.round(0, obj.num_swapping_pairs)  # the reference still attempts the exchange after the last round
        count[-1] = msc.swap_count(reset=True)
    obj.energies_all_runs = None if E_all is None else E_all[:, :num_runs].copy()
    Energy = np.zeros(R)
    obj._EE1_list = []
    for r in range(R):
        EE1 = E_cols[r, :spr].copy()
        Energy[r] = np.min(EE1) if len(EE1) else 0.0
        obj._EE1_list.append(EE1)
    msc.close()
    return M, Energy, count


def apt_preprocessor_chains_production(prob, reps, iter, saved_state, beta, num_sweeps_MCMC, num_sweeps_read,
                                       num_rng):
    """One beta iteration of APT_preprocessor.run (NPT/apt_preprocessor.py:158-179): num_rng independent
    chains in the bit lanes, warm-started from the previous beta's final states (kept on the device)."""
    state = getattr(prob, "_prep_msc", None)
    if state is None or state.n_ladders_requested != num_rng:
        state = _require_msc(prob, [beta], num_rng, _seed_from_numpy())  # random start, as in iteration 1
        prob._prep_msc = state
    else:
        state.set_betas([beta])
    burn = max(0, num_sweeps_MCMC - num_sweeps_read)
    state.sweep(burn)
    Energy = np.zeros((num_rng, min(num_sweeps_read, num_sweeps_MCMC)))
    for t in range(Energy.shape[1]):
        state.sweep(1)
        Energy[:, t] = state.energies()[0, :num_rng]
    return Energy, saved_state


def nmc_run_production(obj, kw):
    raise NotImplementedError("NMC.run has no production mode yet; use mode='replay'")


def apt_icm_run_production(obj, beta_list):
    raise NotImplementedError("APT_ICM.run has no production mode yet; use mode='replay'")
