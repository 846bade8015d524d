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

import os

import numpy as np

from . import _lib, host


def _seed_from_numpy() -> int:
    """A 64-bit Philox seed drawn from the global np.random, so np.random.seed(s) makes runs repeatable."""
    hi, lo = np.random.randint(0, 2**31 - 1, size=2)
    return (int(hi) << 32) | int(lo)


class _NotBitPackable(NotImplementedError):
    """The instance is outside the bit-packed engine (raised by its create call); callers fall back to the generic engines."""


def _require_msc(prob: host.Problem, betas, n_ladders: int, seed: int) -> "_lib.Msc":
    try:
        return _lib.Msc(prob.inst, betas, n_ladders, seed)
    except _lib.NlmcError as e:
        raise _NotBitPackable(
            "the bit-packed engine covers +-J instances with h = 0 and degrees <= 6 (2D/3D lattices, periodic or open, "
            f"Chimera-like graphs) ({e})") from e


def _msc_eligible(prob: host.Problem) -> bool:
    """+-J, h = 0, degrees <= 6: the bit-packed path applies (csrc/nlmc_msc.cu)."""
    cached = getattr(prob, "_msc_eligible", None)
    if cached is None:
        cached = False
        if prob.is_integer and len(prob.val) and not np.any(prob.h):
            deg = np.diff(prob.rp)
            if deg.max() <= 6 and prob.val.max() == 1.0 and prob.val.min() == -1.0:
                import scipy.sparse as sp
                diag = sp.csr_matrix((prob.val, prob.ci, prob.rp), shape=(prob.n, prob.n), copy=False).diagonal()
                cached = bool(np.all(prob.val * prob.val == 1.0)) and not np.any(diag)
        prob._msc_eligible = cached
    return cached


def _engine_energies_exact(prob: host.Problem) -> bool:
    """Integer couplings and fields: the engines' fixed-point / fp32 energies are exact, no fp64 recomputation needed."""
    return bool(prob.is_integer and np.all(prob.h == np.round(prob.h)))


def npt_run_production(obj, beta_list, nmc_kw):
    """NPT.run (NPT/npt.py:535-700) in production mode.  +-J lattices without NMC replicas take the bit-packed
    path; everything else (real-valued or dense J, fields, doNMC replicas) takes the dense tensor-core path."""
    prob = obj._problem()
    if not any(obj.doNMC) and prob.is_integer and getattr(prob, "_msc_eligible", None) is not False:
        # no host-side pre-check of the instance (several passes over the CSR): the create call of the bit-packed engine
        # checks it anyway and reports an instance it cannot hold
        try:
            return _npt_run_msc(obj, prob, beta_list)
        except _NotBitPackable:
            prob._msc_eligible = False
    if _msc_eligible(prob) and not all(obj.doNMC) and not os.environ.get("NLMC_NO_HYBRID"):
        return _npt_run_hybrid(obj, prob, beta_list, nmc_kw)
    return _npt_run_dense(obj, prob, beta_list, nmc_kw)


def _process_group():
    """(dist, world, rank) of an initialised torch.distributed default group with more than one rank, else None."""
    try:
        import torch.distributed as dist
    except Exception:
        return None
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() < 2:
        return None
    return dist, dist.get_world_size(), dist.get_rank()


def _npt_run_msc(obj, prob, beta_list):
    """Bit-packed NPT.  In a multi-rank torch.distributed job (one process per GPU) the temperature range of the
    ladders is sharded over the ranks (distributed.ShardedBetaLadder: only energies cross the GPUs, exchanges permute
    beta labels) and every rank returns the full (M, Energy); otherwise one handle does everything on one GPU."""
    R = obj.num_replicas
    if _process_group() is not None and R >= _process_group()[1] and getattr(obj, "distributed", True):
        return _npt_run_msc_sharded(obj, prob, beta_list)
    spm, spr = obj.num_sweeps_MCMC_per_swap, obj.num_sweeps_read_per_swap
    num_runs = int(getattr(obj, "num_runs", 1))
    n = prob.n
    msc = _require_msc(prob, beta_list[:R], num_runs, _seed_from_numpy())
    count = np.zeros(obj.num_swap_attempts)
    # all rounds but the last run fused on the device; nothing comes back to the host
    for ii in range(obj.num_swap_attempts - 1):
        msc.round(spm, obj.num_swapping_pairs)
    # last round: record the state after every sweep (the reference returns the last round's M); the device writes the
    # record in the layout of M's rows, the host only widens int8 to float64.  (Touching the pages of M from the host threads
    # while the rounds are still running was tried and cost more than it saved: 0.38 s per call against 0.22 s.)
    M = _lib.result_cache.take((R * n, spm)) if spm > 0 else np.zeros((R * n, spm))
    E_cols = np.zeros((R, spm))
    E_all = None
    if spm > 0:
        _, Erec = msc.sweep_record_f64(spm, ladder=0, out=M)  # [R][n][spm] float64: int8 on the wire, widened on arrival
        E_cols[:] = Erec[:, :, 0].T
        E_all = Erec[-1]
    if obj.num_swap_attempts > 0 and spm > 0:
        msc.round(0, obj.num_swapping_pairs)  # the reference still attempts the exchange after the last round
    # accepted exchanges of every round, counted per round on the device (count[ii], NPT/npt.py:664-680); with
    # num_runs ladders side by side the figure is the total over the ladders
    k = min(obj.num_swap_attempts, 4096)
    if k:
        count[-k:] = msc.swap_counts(k)
    obj.energies_all_runs = None if E_all is None else E_all[:, :num_runs].copy()
    Energy = np.zeros(R)
    obj._EE1_list = []
    for r in range(R):
        EE1 = E_cols[r, :spr].copy()
        Energy[r] = np.min(EE1) if len(EE1) else 0.0
        obj._EE1_list.append(EE1)
    msc.close()
    return M, Energy, count


def _npt_run_msc_sharded(obj, prob, beta_list):
    """NPT.run on N GPUs: slots [first, first + count) of every ladder live on this rank.  Per round the ranks
    all-gather one float64 per replica and take identical label decisions.  The last round is recorded on the device;
    the int8 states of run 0 and the energies are all-gathered once (NCCL, device to device) and the ranks listed in
    ``obj.m_on_ranks`` (default: every rank) fetch them and widen them into the reference's float64 M in temperature
    order -- the other ranks return M = None and the same Energy.  With 8 ranks on one host, eight float64 copies of M
    (1 GB each at C5 size) are pure host-memory traffic; ``m_on_ranks = (0,)`` is what a job that post-processes on one
    rank wants."""
    import torch
    from .distributed import ShardedBetaLadder, beta_shard
    dist, world, rank = _process_group()
    R = obj.num_replicas
    spm, spr = obj.num_sweeps_MCMC_per_swap, obj.num_sweeps_read_per_swap
    num_runs = int(getattr(obj, "num_runs", 1))
    n = prob.n
    nccl = dist.get_backend() == "nccl"
    dev = torch.device("cuda", prob.inst.device) if nccl else torch.device("cpu")
    m_ranks = getattr(obj, "m_on_ranks", None)
    want_M = m_ranks is None or rank in tuple(m_ranks)
    seed_t = torch.tensor([_seed_from_numpy() & (2**62 - 1)], dtype=torch.int64, device=dev)
    dist.broadcast(seed_t, 0)  # one seed for the whole job
    try:
        ens = ShardedBetaLadder(prob, beta_list[:R], num_runs, int(seed_t.item()), device=prob.inst.device if nccl else None)
    except _lib.NlmcError as e:
        raise _NotBitPackable(f"mode='production' on several GPUs needs a +-J lattice instance ({e})") from e
    msc = ens.msc
    L = msc.n_ladders
    count = np.zeros(obj.num_swap_attempts)
    for ii in range(obj.num_swap_attempts - 1):
        ens.round(spm, obj.num_swapping_pairs)
    M = np.zeros((R * n, spm)) if want_M else None
    E_cols = np.zeros((R, spm))
    E_all = None
    if spm > 0:
        cmax = beta_shard(R, world, 0)[1]
        shards = [beta_shard(R, world, r) for r in range(world)]
        blk = n * spm                                           # one slot of run 0: a block of rows of M
        if nccl:
            with ens._on_stream():
                pM, pE = msc.sweep_record_dev(spm, ladder=0, rows_of_M=True)      # queued; buffers owned by the handle
                rec_M = torch.as_tensor(_lib.DevArray(pM, (ens.count * blk,), "|i1"), device=dev)
                rec_E = torch.as_tensor(_lib.DevArray(pE, (spm, ens.count, L), "<f8"), device=dev)
                if ens.count == cmax:
                    send_M = rec_M
                else:
                    send_M = torch.zeros(cmax * blk, dtype=torch.int8, device=dev)
                    send_M[:ens.count * blk].copy_(rec_M)
                send_E = torch.zeros((spm, cmax, L), dtype=torch.float64, device=dev)
                send_E[:, :ens.count].copy_(rec_E)
                recv_M = torch.empty(world * cmax * blk, dtype=torch.int8, device=dev)
                recv_E = torch.empty((world, spm, cmax, L), dtype=torch.float64, device=dev)
                dist.all_gather_into_tensor(recv_M, send_M)
                dist.all_gather_into_tensor(recv_E.view(-1, L), send_E.view(-1, L))
                recv_E_h = recv_E.cpu().numpy()                 # waits for the stream
            labels = msc.labels().astype(np.int64)              # labels of the recorded round (the sweeps do not change them)
            if want_M:
                M = _lib.result_cache.take((R * n, spm))
                for r, (f, c) in enumerate(shards):             # slot f+s of run 0 holds temperature labels[f+s, 0]
                    _lib.fetch_widen_blocks(recv_M.data_ptr() + r * cmax * blk, M, c, blk, labels[f:f + c, 0],
                                            device=prob.inst.device, cuda_stream=ens.stream.cuda_stream)
                ens.synchronize()
        else:  # host tensors (gloo): the same data flow without device buffers
            ens.synchronize()
            labels = msc.labels().astype(np.int64)
            Mrec, Erec = msc.sweep_record(spm, ladder=0, energies=True, rows_of_M=True)  # [count][n][spm], [spm][count][ladders]
            send_M = torch.zeros((cmax, n, spm), dtype=torch.int8)
            send_E = torch.zeros((spm, cmax, L), dtype=torch.float64)
            send_M[:ens.count] = torch.from_numpy(Mrec)
            send_E[:, :ens.count] = torch.from_numpy(Erec)
            recv_M = torch.empty((world,) + tuple(send_M.shape), dtype=torch.int8)
            recv_E = torch.empty((world,) + tuple(send_E.shape), dtype=torch.float64)
            dist.all_gather_into_tensor(recv_M.view(-1, spm), send_M.view(-1, spm))
            dist.all_gather_into_tensor(recv_E.view(-1, L), send_E.view(-1, L))
            recv_E_h = recv_E.numpy()
            if want_M:
                Mi8 = np.empty((R, n, spm), dtype=np.int8)
                for r, (f, c) in enumerate(shards):
                    for s in range(c):
                        Mi8[labels[f + s, 0]] = recv_M[r, s].numpy()
                M = _lib.widen_to_f64(Mi8).reshape(R * n, spm)
        E_slots = np.empty((spm, R, L))
        for r, (f, c) in enumerate(shards):
            E_slots[:, f:f + c] = recv_E_h[r, :, :c]
        E_by_beta = np.empty_like(E_slots)
        np.put_along_axis(E_by_beta, np.broadcast_to(labels[None], E_slots.shape), E_slots, axis=1)
        E_cols[:] = E_by_beta[:, :, 0].T
        E_all = E_by_beta[-1]
    if obj.num_swap_attempts > 0 and spm > 0:
        ens.round(0, obj.num_swapping_pairs)
    ens.synchronize()
    k = min(obj.num_swap_attempts, 4096)
    if k:
        count[-k:] = msc.swap_counts(k)
    obj.energies_all_runs = None if E_all is None else E_all[:, :num_runs].copy()
    Energy = np.zeros(R)
    obj._EE1_list = []
    for r in range(R):
        EE1 = E_cols[r, :spr].copy()
        Energy[r] = np.min(EE1) if len(EE1) else 0.0
        obj._EE1_list.append(EE1)
    ens.close()
    return M, Energy, count


# ---------------------------------------------------------------------------------------------------
# generic engines: sparse graph-coloured (K2a) or dense tensor-core (K3); one row per replica, beta per row,
# NMC phases as per-site modes
# ---------------------------------------------------------------------------------------------------
def _generic_engine(prob: host.Problem, betas, seed: int):
    """Sparse instances (mean degree <= 128) take the graph-coloured kernel (state in shared memory, or in a global
    workspace when it does not fit); dense ones (SK) the tensor-core path, where a colouring would degenerate to one
    site per colour."""
    betas = np.asarray(betas, dtype=np.float64)
    mean_degree = len(prob.val) / max(prob.n, 1)
    if mean_degree <= 128 and not os.environ.get("NLMC_FORCE_DENSE"):
        try:
            return _lib.Col(prob.inst, betas, seed=seed)
        except _lib.NlmcError:
            pass
    return _lib.Dense(prob.inst, betas, n_split=3, seed=seed)


def _backbones(prob, states, global_beta, nmc_kw):
    """LBP backbone (K5 + host lambda schedule) of every state in `states` [G][n] -> list of index arrays."""
    from .nmc_core import lbp_convexified
    lbp = getattr(prob, "_lbp_handle", None)  # reverse-entry index and summation programs are built once per instance
    if lbp is None:
        lbp = prob._lbp_handle = _lib.Lbp(prob.lbp_instance())
    out = []
    for m_star in states:
        out.append(lbp_convexified(prob, lbp, m_star.astype(np.float64), nmc_kw["lambda_start"], nmc_kw["lambda_end"],
                                   nmc_kw["lambda_reduction_factor"], nmc_kw["tolerance"], nmc_kw["max_iterations"],
                                   nmc_kw["threshold_initial"], nmc_kw["threshold_cutoff"], global_beta,
                                   as_index_set=True))
    return out


def _nmc_cycles_dense(prob, d, m_star, nmc_kw, variant, record_run0=True, all_clusters=None, keep_states=True):
    """NMC_subroutine (NMC/nmc.py:320-440 / NPT/npt.py:357-477) for all rows of the engine handle `d` in lock step;
    rows are independent chains.  Returns per row (M_overall [n][cols], energy_overall [cols], clusters).  With
    keep_states=False only the energies are recorded and M_overall holds the final state as its single column (the
    rounds of NPT.run before the last one need nothing else)."""
    G, n = d.R, prob.n
    num_cycles, phase = nmc_kw["num_cycles"], nmc_kw["phase_sweeps"]
    fuf, M_skip, temp_x = nmc_kw["full_update_frequency"], nmc_kw["M_skip"], nmc_kw["temp_x"]
    m_init = np.asarray(m_star, dtype=np.int8).reshape(G, n).copy()
    m_star = m_init.copy()
    cols = [[] for _ in range(G)]   # recorded states per row (every M_skip-th sweep of every phase)
    ens = [[] for _ in range(G)]
    clusters = None

    def run_phase(modes):
        nonlocal m_init
        d.set_spins(m_init)
        d.set_site_modes(modes, temp_x)
        d.best_reset()
        if hasattr(d, "sweep_record"):  # K2a: the whole phase is one launch, recording and argmin on the device
            states, E = d.sweep_record(phase, record_every=M_skip, track_best=True, want_states=keep_states)
            for g in range(G):
                if keep_states:
                    cols[g].extend(states[:, g])
                ens[g].extend(E[::M_skip, g])
            m_init, _ = d.best_get()
            return
        for j in range(phase):
            d.sweep(1)
            E = d.best_update()
            if j % M_skip == 0:
                S = d.get_spins() if keep_states else None
                for g in range(G):
                    if keep_states:
                        cols[g].append(S[g].copy())
                    ens[g].append(E[g])
        m_init, _ = d.best_get()  # m_init = M[:, argmin E], first minimum wins (nmc.py:394-395)

    if all_clusters is not None:  # clusters provided by the caller: no LBP (nmc.py:357,367)
        clusters = [np.asarray(all_clusters, dtype=int)] * G
    elif variant == "npt":
        clusters = _backbones(prob, m_star, nmc_kw["global_beta"], nmc_kw)
    for cycle in range(num_cycles):
        if variant == "nmc" and all_clusters is None:
            clusters = _backbones(prob, m_star, nmc_kw["global_beta"], nmc_kw)
        in_cl = np.zeros((G, n), dtype=bool)
        for g in range(G):
            in_cl[g, clusters[g]] = True
        run_phase(np.where(in_cl, 1, 2).astype(np.uint8))   # C: backbone hot, the rest frozen (nmc.py:377-385)
        run_phase(np.where(in_cl, 2, 0).astype(np.uint8))   # NC: backbone frozen (nmc.py:398-406)
        if cycle % fuf == 0:
            run_phase(None)                                  # ALL (nmc.py:419-433)
            if variant == "nmc":
                m_star = m_init.copy()
    d.set_site_modes(None)
    out = []
    final = None if keep_states else d.get_spins()
    for g in range(G):
        if not keep_states:
            out.append((final[g].astype(np.float64)[:, None], np.array(ens[g], dtype=np.float64),
                        clusters[g] if clusters else np.array([], dtype=int)))
            continue
        Mo = np.array(cols[g], dtype=np.float64).T if cols[g] else np.zeros((n, 0))
        Eo = np.array(ens[g], dtype=np.float64)
        if cols[g] and not _engine_energies_exact(prob):
            # the engines' own energies are fixed-point (K2a) or fp32 (K3): good enough to steer the run, but the
            # RETURNED energies are those of the returned states in fp64 (kernel K4), as the reference's are
            Eo = prob.inst.energy_states(np.array(cols[g], dtype=np.int8))
        out.append((Mo, Eo, clusters[g] if clusters else np.array([], dtype=int)))
    return out


def _npt_run_generic_labels(obj, prob, beta_list):
    """NPT.run without NMC replicas on the generic engines (K2a sparse / K3 dense): sweeps, energies and the replica
    exchange stay on the device for every round -- the exchange permutes beta labels (nlmc_col_exchange /
    nlmc_dense_exchange), no spin leaves the GPU until the last round's states are recorded.  `num_runs` independent
    ladders ride as further rows (row = run * R + slot); the reference's tuple is that of run 0."""
    R, n = obj.num_replicas, prob.n
    spm, spr = obj.num_sweeps_MCMC_per_swap, obj.num_sweeps_read_per_swap
    runs = max(1, int(getattr(obj, "num_runs", 1)))
    betas = np.asarray(beta_list[:R], dtype=np.float64)
    d = _generic_engine(prob, np.tile(betas, runs), _seed_from_numpy())
    d.ladders(betas)
    d.set_spins(np.sign(2 * np.random.rand(runs * R, n) - 1).astype(np.int8))  # NPT/npt.py:612, once per run
    count = np.zeros(obj.num_swap_attempts)
    for ii in range(obj.num_swap_attempts - 1):
        d.sweep(spm)
        d.exchange(obj.num_swapping_pairs)
    M = np.zeros((R * n, spm))
    E_cols = np.zeros((R, spm))
    E_all = None
    if spm > 0:
        labels, _ = d.labels(0)                     # fixed during the round's sweeps
        lab = labels.reshape(runs, R)
        if hasattr(d, "sweep_record"):
            states, E = d.sweep_record(spm, record_every=1)          # [spm][rows][n], [spm][rows]
            states0 = states[:, :R]
        else:
            states0 = np.empty((spm, R, n), dtype=np.int8)
            E = np.empty((spm, runs * R))
            for j in range(spm):
                d.sweep(1)
                E[j] = d.energies()
                states0[j] = d.get_spins()[:R]
        if not _engine_energies_exact(prob):   # returned energies in fp64 from the returned states (K4)
            E = E.copy()
            E[:, :R] = prob.inst.energy_states(np.ascontiguousarray(states0).reshape(-1, n)).reshape(spm, R)
        M3 = M.reshape(R, n, spm)
        for s in range(R):                          # slot s of run 0 holds temperature index lab[0, s]
            M3[lab[0, s]] = states0[:, s, :].T
            E_cols[lab[0, s]] = E[:, s]
        E_last = E[-1].reshape(runs, R)
        E_all = np.empty((R, runs))
        for r in range(runs):
            E_all[lab[r], r] = E_last[r]
    if obj.num_swap_attempts > 0 and spm > 0:
        d.exchange(obj.num_swapping_pairs)  # the reference still attempts the exchange after the last round
    k = min(obj.num_swap_attempts, 4096)
    if k:
        count[-k:] = d.labels(k)[1]
    obj.energies_all_runs = E_all
    Energy = np.zeros(R)
    obj._EE1_list = []
    for r in range(R):
        EE1 = E_cols[r, :spr].copy()
        Energy[r] = np.min(EE1) if len(EE1) else 0.0
        obj._EE1_list.append(EE1)
    d.close()
    return M, Energy, count


def _npt_run_hybrid(obj, prob, beta_list, nmc_kw):
    """NPT.run on a +-J lattice with NMC on some replicas (config C2 as stated): the plain replicas stay on the bit-packed
    engine (K2, one ladder in the bit lanes), only the doNMC replicas run on the graph-coloured engine (K2a, per-site
    phase modes) at global_beta.  The two kinds sweep concurrently on their own streams.  One set of pairs is drawn per
    round as in the reference (NPT/npt.py:649-680, host RNG); an accepted exchange moves the two configurations -- between
    two slots of the bit-packed handle, or between a slot and an NMC row (2 x n bytes through the host; the kinds differ,
    so a label cannot stand in for the configuration)."""
    R, n = obj.num_replicas, prob.n
    spm, spr = obj.num_sweeps_MCMC_per_swap, obj.num_sweeps_read_per_swap
    mc_ids = [r for r in range(R) if not obj.doNMC[r]]
    nmc_ids = [r for r in range(R) if obj.doNMC[r]]
    slot = {r: k for k, r in enumerate(mc_ids)}      # replica -> slot of the bit-packed handle
    row = {r: k for k, r in enumerate(nmc_ids)}      # replica -> row of the NMC handle
    seed = _seed_from_numpy()
    msc = _require_msc(prob, np.asarray(beta_list, dtype=np.float64)[mc_ids], 1, seed)
    d_nmc = _generic_engine(prob, np.full(len(nmc_ids), float(nmc_kw["global_beta"])), seed + 1)
    state = np.sign(2 * np.random.rand(R, n) - 1).astype(np.int8)  # NPT/npt.py:612
    for r in mc_ids:
        msc.set_spins(slot[r], 0, state[r])
    M = np.empty((R * n, spm))
    M3 = M.reshape(R, n, spm)
    E_cols = np.zeros((R, spm))
    E_last = np.zeros(R)
    count = np.zeros(obj.num_swap_attempts)
    all_pairs = [(i, i + 1) for i in range(1, R)]
    for ii in range(obj.num_swap_attempts):
        last = ii == obj.num_swap_attempts - 1
        if not last:
            msc.sweep(spm)                       # queued on the handle's stream; the NMC cycles below run meanwhile
        res = _nmc_cycles_dense(prob, d_nmc, state[nmc_ids], nmc_kw, "npt", keep_states=last)
        for g, r in enumerate(nmc_ids):
            Mo, Eo, _ = res[g]
            if last:
                M3[r] = Mo[:, -spm:]
                E_cols[r] = Eo[-spm:]
            E_last[r] = Eo[-1]
            state[r] = Mo[:, -1].astype(np.int8)
        if last and spm > 0:
            Mrec, Erec = msc.sweep_record(spm, ladder=0, energies=True, rows_of_M=True)   # [slots][n][spm]
            for r in mc_ids:
                _lib.widen_to_f64(Mrec[slot[r]], out=M3[r])
                E_cols[r] = Erec[:, slot[r], 0]
            E_mc = Erec[-1][:, 0]
        else:
            E_mc = msc.energies()[:, 0]
        for r in mc_ids:
            E_last[r] = E_mc[slot[r]]
        for sel, nxt in host.select_non_overlapping_pairs(all_pairs, obj.num_swapping_pairs):
            a, b = sel - 1, nxt - 1
            dE = E_last[b] - E_last[a]
            dB = beta_list[b] - beta_list[a]
            if np.random.rand() < min(1, np.exp(dB * dE)):
                count[ii] += 1
                if last:
                    continue                     # nothing reads m_start after the last round
                ca = msc.get_spins(slot[a], 0) if a in slot else state[a].copy()
                cb = msc.get_spins(slot[b], 0) if b in slot else state[b].copy()
                for r, c in ((a, cb), (b, ca)):
                    if r in slot:
                        msc.set_spins(slot[r], 0, c)
                    else:
                        state[r] = c
    Energy = np.zeros(R)
    obj._EE1_list = []
    for r in range(R):
        EE1 = E_cols[r, :spr].copy()
        Energy[r] = np.min(EE1) if len(EE1) else 0.0
        obj._EE1_list.append(EE1)
    obj.energies_all_runs = None
    msc.close()
    d_nmc.close()
    return M, Energy, count


def _npt_run_dense(obj, prob, beta_list, nmc_kw):
    """Generic-engine NPT.  Without NMC replicas everything stays on the device (_npt_run_generic_labels).  With them,
    plain replicas and doNMC replicas live in two handles (their sweep counts per round differ, NPT/npt.py:577-580) and an
    exchange between the two kinds has to move the 2 x n spins of the accepted pair, which is done through the host."""
    if not any(obj.doNMC):
        return _npt_run_generic_labels(obj, prob, beta_list)
    R, n = obj.num_replicas, prob.n
    spm, spr = obj.num_sweeps_MCMC_per_swap, obj.num_sweeps_read_per_swap
    mc_ids = [r for r in range(R) if not obj.doNMC[r]]
    nmc_ids = [r for r in range(R) if obj.doNMC[r]]
    seed = _seed_from_numpy()
    d_mc = _generic_engine(prob, beta_list[mc_ids], seed) if mc_ids else None
    d_nmc = _generic_engine(prob, np.full(len(nmc_ids), float(nmc_kw["global_beta"])), seed + 1) \
        if nmc_ids else None  # doNMC replicas run at global_beta, not beta_list[i] (NPT/npt.py:630-637)
    state = np.sign(2 * np.random.rand(R, n) - 1).astype(np.int8)  # NPT/npt.py:612
    M = np.zeros((R * n, spm))
    E_cols = np.zeros((R, spm))
    count = np.zeros(obj.num_swap_attempts)
    all_pairs = [(i, i + 1) for i in range(1, R)]
    for ii in range(obj.num_swap_attempts):
        last = ii == obj.num_swap_attempts - 1
        if mc_ids:
            d_mc.set_spins(state[mc_ids])
            if last and hasattr(d_mc, "sweep_record"):
                states, E = d_mc.sweep_record(spm, record_every=1)
                if not _engine_energies_exact(prob):  # returned energies in fp64 from the returned states (K4)
                    E = prob.inst.energy_states(states.reshape(-1, n)).reshape(spm, len(mc_ids))
                M.reshape(R, n, spm)[mc_ids] = np.ascontiguousarray(states.transpose(1, 2, 0))  # blocked transpose in int8
                E_cols[mc_ids] = E.T
            elif last:
                for j in range(spm):
                    d_mc.sweep(1)
                    E = d_mc.energies()
                    S = d_mc.get_spins()
                    if not _engine_energies_exact(prob):
                        E = prob.inst.energy_states(S)
                    for g, r in enumerate(mc_ids):
                        M[r * n:(r + 1) * n, j] = S[g]
                        E_cols[r, j] = E[g]
            else:
                d_mc.sweep(spm)
                E_cols[mc_ids, -1] = d_mc.energies()
            state[mc_ids] = d_mc.get_spins()
        if nmc_ids:
            res = _nmc_cycles_dense(prob, d_nmc, state[nmc_ids], nmc_kw, "npt", keep_states=last)
            for g, r in enumerate(nmc_ids):
                Mo, Eo, _ = res[g]
                if last:
                    M[r * n:(r + 1) * n, :] = Mo[:, -spm:]
                E_cols[r] = Eo[-spm:]
                state[r] = Mo[:, -1].astype(np.int8)
        for sel, nxt in host.select_non_overlapping_pairs(all_pairs, obj.num_swapping_pairs):
            dE = E_cols[nxt - 1, -1] - E_cols[sel - 1, -1]
            dB = beta_list[nxt - 1] - beta_list[sel - 1]
            if np.random.rand() < min(1, np.exp(dB * dE)):
                count[ii] += 1
                state[[sel - 1, nxt - 1]] = state[[nxt - 1, sel - 1]]  # only m_start is exchanged (npt.py:677-678)
    Energy = np.zeros(R)
    obj._EE1_list = []
    for r in range(R):
        EE1 = E_cols[r, :spr].copy()
        Energy[r] = np.min(EE1) if len(EE1) else 0.0
        obj._EE1_list.append(EE1)
    obj.energies_all_runs = None
    for d in (d_mc, d_nmc):
        if d is not None:
            d.close()
    return M, Energy, count


def nmc_run_production(obj, kw):
    """NMC.run (NMC/nmc.py:442-520) on the dense engine: annealed MCMC from beta 0 to global_beta with the best
    state tracked on the device, then NMC cycles (LBP backbone K5, hot-backbone / frozen phases as site modes)."""
    prob = host.Problem(obj.J, obj.h, obj.device)
    n = prob.n
    global_beta = kw["global_beta"]
    d = _generic_engine(prob, [float(global_beta)], _seed_from_numpy())
    d.set_spins(np.sign(2 * np.random.rand(n) - 1).astype(np.int8)[None, :])  # nmc.py:487
    sched = host.beta_schedule(kw["num_sweeps_initial"], global_beta, anneal=True, sweeps_per_beta=1, initial_beta=0)
    d.best_reset()
    if hasattr(d, "sweep_record"):  # the whole annealing leg is one launch
        if len(sched):
            d.sweep_record(len(sched), track_best=True, beta_sched=sched[:, None], want_states=False, want_energies=False)
    else:
        for b in sched:
            d.set_betas([max(float(b), 1e-12)])
            d.sweep(1)
            d.best_update(fetch=False)
        d.set_betas([float(global_beta)])
    m_star, E_star = d.best_get()
    if obj.verbose:
        print(f'\ninitial m_star energy = {E_star[0]:.8f}')
    nmc_kw = dict(num_cycles=kw["num_NMC_cycles"], phase_sweeps=kw["num_sweeps_per_NMC_phase"],
                  full_update_frequency=kw["full_update_frequency"], M_skip=kw["M_skip"], global_beta=global_beta,
                  temp_x=kw["temp_x"], lambda_start=kw["lambda_start"], lambda_end=kw["lambda_end"],
                  lambda_reduction_factor=kw["lambda_reduction_factor"], threshold_initial=kw["threshold_initial"],
                  threshold_cutoff=kw["threshold_cutoff"], max_iterations=kw["max_iterations"], tolerance=kw["tolerance"])
    Mo, Eo, clusters = _nmc_cycles_dense(prob, d, m_star, nmc_kw, "nmc")[0]
    d.close()
    obj.all_clusters = clusters
    return Mo, Eo, float(np.min(Eo))


def nmc_subroutine_production(obj, prob, m_star, phase_sweeps, kw, variant, all_clusters=None):
    """The public NMC_subroutine method in production mode: one chain on the generic engine.  Returns the
    reference's tuple (M_overall, energy_overall, min_energy, all_clusters)."""
    d = _generic_engine(prob, [float(kw["global_beta"])], _seed_from_numpy())
    try:
        nmc_kw = dict(kw, phase_sweeps=phase_sweeps)
        Mo, Eo, clusters = _nmc_cycles_dense(prob, d, np.asarray(m_star).reshape(1, -1), nmc_kw, variant,
                                             all_clusters=all_clusters)[0]
    finally:
        d.close()
    return Mo, Eo, (float(np.min(Eo)) if len(Eo) else 0.0), clusters


def apt_preprocessor_chains_production(prob, reps, iter, saved_state, beta, num_sweeps_MCMC, num_sweeps_read,
                                       num_rng):
    """One beta iteration of APT_preprocessor.run (NPT/apt_preprocessor.py:158-179): num_rng independent chains,
    warm-started from the previous beta's final states (kept on the device).  +-J lattices ride in the bit lanes (K2);
    every other instance (real-valued J, fields, any degree) runs one chain per row of the generic engine (K2a / K3)."""
    burn = max(0, num_sweeps_MCMC - num_sweeps_read)
    n_read = min(num_sweeps_read, num_sweeps_MCMC)
    if not _msc_eligible(prob):
        state = getattr(prob, "_prep_gen", None)
        if state is None or state.R != num_rng:
            state = _generic_engine(prob, np.full(num_rng, float(beta)), _seed_from_numpy())
            state.set_spins(np.sign(2. * np.random.rand(num_rng, prob.n) - 1).astype(np.int8))  # apt_preprocessor.py:164
            prob._prep_gen = state
        else:
            state.set_betas(np.full(num_rng, float(beta)))
        if burn:
            state.sweep(burn)
        if n_read == 0:
            return np.zeros((num_rng, 0)), saved_state
        if hasattr(state, "sweep_record"):  # K2a: per-sweep energies recorded inside the launch
            _, E = state.sweep_record(n_read, want_states=False, want_energies=True)
            return np.ascontiguousarray(E.T), saved_state
        Energy = np.empty((num_rng, n_read))
        for j in range(n_read):
            state.sweep(1)
            Energy[:, j] = state.energies()
        return Energy, saved_state
    state = getattr(prob, "_prep_msc", None)
    if state is None or state.n_ladders_requested != num_rng:
        state = _require_msc(prob, [beta], num_rng, _seed_from_numpy())  # random start, as in iteration 1
        prob._prep_msc = state
    else:
        state.set_betas([beta])
    state.sweep(burn)
    _, Erec = state.sweep_record(n_read, ladder=None, energies=True)  # [n_read][1][ladders], recorded on the device
    Energy = np.ascontiguousarray(Erec[:, 0, :num_rng].T) if n_read else np.zeros((num_rng, 0))
    return Energy, saved_state


class _IcmEngine:
    """The R x 10 chains of APT_ICM on whichever production engine fits the instance: the bit-packed path
    (sub-replica s of beta b = lane s of the words of beta b) or the dense tensor-core path (row b*S + s)."""

    def __init__(self, prob, betas, S, seed):
        self.R, self.S, self.n = len(betas), S, prob.n
        self.msc = _msc_eligible(prob)
        if self.msc:
            self.eng = _lib.Msc(prob.inst, betas, S, seed)
        else:
            self.eng = _generic_engine(prob, np.repeat(betas, S), seed)

    def sweep(self, k):
        self.eng.sweep(k)

    def device_round(self, spm, num_pairs):
        """spm sweeps, energies and the exchange entirely on the device (K2 + K4' + K6); returns the number of
        accepted exchanges among the S sub-replicas, or None when the engine has no device-side exchange.  K6 draws
        the non-overlapping pairs per chain (the reference draws one set per round and tests it for each of the 10
        sub-replicas, apt_ICM.py:249-285): every chain sees the same selection and acceptance law."""
        if not self.msc:
            return None
        self.eng.swap_count(reset=True)
        self.eng.round(spm, num_pairs)
        # the handle pads the S chains to 128 lanes; report the share of the S real ones
        return self.eng.swap_count(reset=True) * self.S / self.eng.n_ladders

    def energies(self):  # [R][S]
        E = self.eng.energies()
        return E[:, :self.S].copy() if self.msc else E.reshape(self.R, self.S)

    def get_states(self):  # int8 [R][S][n]
        if self.msc:
            return np.stack([np.stack([self.eng.get_spins(b, s) for s in range(self.S)]) for b in range(self.R)])
        return self.eng.get_spins().reshape(self.R, self.S, self.n)

    def set_states(self, X):
        if self.msc:
            for b in range(self.R):
                for s in range(self.S):
                    self.eng.set_spins(b, s, X[b, s])
        else:
            self.eng.set_spins(X.reshape(self.R * self.S, self.n))

    def close(self):
        self.eng.close()


def apt_icm_run_production(obj, beta_list):
    """APT_ICM.run (NPT/apt_ICM.py:145-305) in production mode, reference semantics: the Houdayer move edits the
    FIRST column of each sub-replica's block of M while the exchange reads the LAST one and the chains continue
    from the unedited states, so the edits feed back only when there is one sweep per swap AND the exchange
    of that pair is accepted (apt_ICM.py:213,282-283)."""
    S = 10  # num_subreplicas, apt_ICM.py:177
    R, spm, spr = obj.num_replicas, obj.num_sweeps_MCMC_per_swap, obj.num_sweeps_read_per_swap
    if spm < 1:
        raise ValueError("num_sweeps_MCMC must be at least num_swap_attempts")
    prob = host.Problem(obj.J, obj.h, obj.device)
    n = prob.n
    eng = _IcmEngine(prob, beta_list[:R], S, _seed_from_numpy())
    all_pairs = [(i, i + 1) for i in range(1, R)]
    M = np.zeros((n * R, spm * S))
    E_cols = np.zeros((R, S, spm))
    count = np.zeros(obj.num_swap_attempts)
    for ii in range(int(obj.num_swap_attempts)):
        last_round = ii == obj.num_swap_attempts - 1
        need_first = last_round or spm == 1   # the Houdayer edit is only observable in these cases
        first = E_first = None
        cols, ens = [], []
        if not need_first:
            acc = eng.device_round(spm, obj.num_swapping_pairs)
            if acc is not None:  # nothing of this round is observable on the host: it stays on the device
                count[ii] = acc
                continue
        if need_first:
            eng.sweep(1)
            first, E_first = eng.get_states(), eng.energies()
            cols, ens = [first], [E_first]
            for j in range(1, spm):
                eng.sweep(1)
                if last_round:  # every column of the returned M
                    cols.append(eng.get_states())
                    ens.append(eng.energies())
        else:
            eng.sweep(spm)
        E_last = eng.energies() if spm > 1 else None
        unedited = first.copy() if (need_first and spm == 1 and not last_round) else None
        if need_first:  # Houdayer move on the first column (apt_ICM.py:216-246)
            for r in range(R):
                shuffled = np.random.permutation(S)
                pairs = [(int(shuffled[2 * p]), int(shuffled[2 * p + 1])) for p in range(S // 2)]
                s1 = np.stack([first[r, a] for a, _ in pairs])
                s2 = np.stack([first[r, b] for _, b in pairs])
                labels, counts = _lib.icm_clusters(prob.inst, s1, s2)
                for p, (a, b) in enumerate(pairs):
                    if not counts[p]:
                        continue
                    members = labels[p] == np.random.randint(int(counts[p]))
                    if int(members.sum()) > n // 2:        # Katzgraber rule (apt_ICM.py:236-237)
                        first[r, a] = -first[r, a]
                    else:
                        first[r, a][members], first[r, b][members] = s2[p][members], s1[p][members]
            E_first = prob.inst.energy_states(first.reshape(R * S, n)).reshape(R, S)
            ens[0] = E_first
            if spm == 1:
                E_last = E_first
        if last_round:
            for j in range(spm):
                for s in range(S):
                    M[:, s * spm + j] = cols[j][:, s, :].reshape(-1)
                E_cols[:, :, j] = ens[j]
        # exchange per sub-replica on the last column (apt_ICM.py:251-285); the chains continue from it
        accepted = []
        for sel, nxt in host.select_non_overlapping_pairs(all_pairs, obj.num_swapping_pairs):
            dB = beta_list[nxt - 1] - beta_list[sel - 1]
            for s in range(S):
                dE = E_last[nxt - 1, s] - E_last[sel - 1, s]
                if np.random.rand() < min(1, np.exp(dB * dE)):
                    accepted.append((sel - 1, nxt - 1, s))
        count[ii] = len(accepted)
        if not last_round and accepted:
            if spm == 1:
                # one sweep per swap: the exchange reads the (edited) first column, but m_start_matrix was taken before the
                # edit (apt_ICM.py:213) -- only the accepted pairs receive edited states (apt_ICM.py:282-283), every other
                # chain continues from its unedited state
                X = unedited
                for a, b, s in accepted:
                    X[a, s], X[b, s] = first[b, s].copy(), first[a, s].copy()
            else:
                X = eng.get_states()
                for a, b, s in accepted:
                    X[[a, b], s] = X[[b, a], s]
            eng.set_states(X)
    eng.close()
    Energy = np.zeros(R)
    E_flat = E_cols.reshape(R, S * spm)
    for r in range(R):
        Energy[r] = np.min(E_flat[r, :spr]) if spr else 0.0
    return M, Energy
