#!/usr/bin/env python
"""bench.py -- headline benchmark of nlmc_b200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Metric (BASELINE.json): spin-flip attempts/s.  Workload: config C5 as stated -- 3D +-J Edwards-Anderson L=64
(262,144 spins), NPT with 32 inverse temperatures x 128 independent ladders = 4096 replicas IN TOTAL.  A "step" is one
swap round of NPT.run for all ladders: `spm` heat-bath sweeps of every replica, the energies of every replica, and the
replica-exchange step (NPT/npt.py:617-680).

  value      attempts/s with the state resident in HBM, CUDA events on the stream the kernels run on, max over ranks.
             N > 1 (torchrun, one process per GPU): STRONG scaling -- the temperature range of the 128 ladders is sharded
             over the GPUs; per round one float64 per replica is all-gathered over NCCL and every rank permutes the beta
             labels identically (spins never move).  `weak` holds the weak-scaling figure beside it (every GPU runs the
             whole stated C5 on its own ladders, no collective).
  e2e        the same metric through the drop-in class, NPT(J, h, mode="production").run(...): host J in, numpy M and
             Energy out, one call = K rounds; `e2e_host_buffers` is the C-ABI round that ships the packed state both ways.
  roofline   the colour-sweep kernel against the measured HBM peak with the bit-packed layout's algorithmic bytes
             (0.25 B/attempt); `roofline_issue` is the bound that applies (integer issue).
  sustained  the same step for >= 5 s with clocks and power sampled under load.
  time_to_target  median time to the shipped ground-state energy (Chimera droplet, DCL, Wishart), GPU vs the CPU port.
  --impl reference   the reference algorithm on the host cores: the oracle C port of MCMC (oracle/nlmc_oracle.c; the
             reference itself is pure Python and cannot travel to the GPU box), all host threads, a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "nonlocal-monte-carlo_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "spin_flip_attempts_per_s"
UNIT = "attempts/s"
SURVEY_BYTES_PER_ATTEMPT = 42.0  # B_alg = 6 + 6*deg for +-J int8 CSR, deg = 6


def ea3d_csr(L: int, seed: int):
    """3D periodic +-J EA instance of SURVEY.md 8(d) as scipy CSR (site i = x + L*(y + L*z))."""
    from nlmc_b200 import instances
    return instances.ea3d_pm_j(L, seed)[0]


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={device_index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.2] or [r for _, r in self.rows]
        try:
            sm = [float(r[0]) for r in rows]
            reasons = []
            for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6)):
                if any(r[col].lower().startswith("active") for r in rows):
                    reasons.append(name)
            return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(rows[0][1]), "reasons": reasons,
                    "power_w_max": max(float(r[2]) for r in rows), "samples": len(rows)}
        except Exception:
            return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return None


# ---------------------------------------------------------------------------------------------------
def cpu_port_rate(L: int, betas, reps: int, sweeps: int, threads: int, seed: int = 5):
    """attempts/s of the oracle C port (reference algorithm: random-permutation sequential heat bath with an
    injected MT19937 stream) on `threads` host threads; one call = reps x sweeps sweeps of an L^3 lattice."""
    from oracle import oracle as O
    A = ea3d_csr(L, seed)
    csr = O.Csr(A)
    n = csr.n
    rs = np.random.RandomState(1)
    perm = np.empty((reps, sweeps, n), dtype=np.int32)
    u = np.empty((reps, sweeps, n), dtype=np.float64)
    for r in range(reps):
        for s in range(sweeps):
            perm[r, s] = rs.permutation(n)
        u[r] = rs.rand(sweeps, n)
    m = rs.choice([-1, 1], size=(reps, n)).astype(np.int8)
    beta_run = np.repeat(np.resize(np.asarray(betas, dtype=np.float64), reps)[:, None], sweeps, axis=1).copy()
    h = np.zeros(n)

    def once():
        t0 = time.perf_counter()
        used = O.lib().nlmc_oracle_mcmc_many(reps, n, csr.rp, csr.ci, csr.val, h, sweeps, beta_run, perm, u, m, threads)
        return time.perf_counter() - t0, used

    return once, reps * sweeps * n


def run_reference(args, rank, world):
    if rank != 0:
        return
    L = args.L
    cores = os.cpu_count() or 1
    reps = max(1, min(cores, 64))
    sweeps = max(1, args.ref_sweeps)
    betas = np.linspace(0.2, 2.0, args.n_beta)
    once, attempts = cpu_port_rate(L, betas, reps, sweeps, cores)
    for _ in range(args.warmup):
        once()
    t = 0.0
    used = cores
    for _ in range(args.steps):
        dt, used = once()
        t += dt
    value = attempts * args.steps / t
    sample = f"{reps} replicas x {sweeps} sweeps of the L={L} lattice per step ({attempts:.3g} attempts)"
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": workload_config(args, world, reference=True),
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def workload_config(args, world, reference=False):
    lad = ((args.n_ladders + 127) // 128) * 128
    return {"workload": f"C5: 3D +-J Edwards-Anderson L={args.L} ({args.L ** 3} spins) NPT, {args.n_beta} betas x "
                        f"{lad} ladders = {args.n_beta * lad} replicas in total, "
                        f"{args.spm} sweeps per swap round, {args.pairs} swapping pairs per ladder",
            "L": args.L, "n_beta": args.n_beta, "n_ladders": lad, "sweeps_per_step": args.spm,
            "replicas_total": args.n_beta * lad, "beta_range": [0.2, 2.0],
            "parallelism": ("one GPU owns every temperature slot" if world == 1 else
                            f"temperature range of every ladder sharded over {world} GPUs in contiguous blocks; per round one "
                            "float64 energy per replica is all-gathered (NCCL) and every rank permutes the beta labels "
                            "identically; spin configurations never move") if not reference else "host threads",
            "l2": "state (n x words x 4 B = 134 MB at the default size) is larger than the 126 MB L2 on one GPU; a block of a "
                  "sharded ladder (134 MB / N) fits, which is part of what the per-N numbers show; no flush"
            if not reference else "n/a (host)"}


def bind_to_gpu_numa_node(device_index: int):
    """Pin this rank to the CPUs of its GPU's NUMA node (pinned host buffers are first-touched next to the GPU)."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(device_index)
        path = f"/sys/bus/pci/devices/{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


def load_ttt():
    import importlib.util
    spec = importlib.util.spec_from_file_location("time_to_target", os.path.join(ROOT, "tools", "time_to_target.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run_ours(args, rank, world, local_rank):
    import torch
    from nlmc_b200 import NPT, _lib, host
    from nlmc_b200.distributed import ShardedBetaLadder
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    device = local_rank
    torch.cuda.set_device(device)
    _lib.require_device(device)
    numa_node = bind_to_gpu_numa_node(device) if world > 1 else None
    L, n_beta, spm, pairs = args.L, args.n_beta, args.spm, args.pairs
    betas = np.linspace(0.2, 2.0, n_beta)
    sampler = ClockSampler(device) if rank == 0 else None  # started early: nvidia-smi needs ~0.1 s before its first row
    A = ea3d_csr(L, 5)
    n = A.shape[0]
    prob = host.Problem(A, np.zeros(n), device=device)
    stream = torch.cuda.Stream(device)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(device)

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], device=f"cuda:{device}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(step, steps):
        """ms for `steps` calls of step(), CUDA events on the stream the kernels are launched on, max over ranks."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        w0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        e1.synchronize()
        w1 = time.perf_counter()
        ms = e0.elapsed_time(e1)
        barrier()
        return max_over_ranks(ms), w0, w1

    # ---- headline: the stated C5 (n_beta x n_ladders replicas IN TOTAL), state resident in HBM ----------------------
    if world == 1:
        eng = _lib.Msc(prob.inst, betas, args.n_ladders, seed=1000)
        eng.set_stream(stream.cuda_stream)
        msc, step = eng, (lambda: eng.round(spm, pairs))
        exchange = "configuration bits exchanged on the device (msc_swap_decide + msc_swap_apply)"
        kernels_per_step = spm * (eng.n_colours + 1) + 2 + 4   # sweeps + counter bumps, energy (2), exchange (4)
    else:
        ens = ShardedBetaLadder(prob, betas, args.n_ladders, seed=1000, device=device)
        stream = ens.stream
        msc, step = ens.msc, (lambda: ens.round(spm, pairs))
        exchange = ("beta labels: energies all-gathered over NCCL (8 B per replica), identical Philox-keyed decisions on every "
                    "rank, thresholds rebuilt from the labels (msc_label_swap + msc_thrbits)")
        bumps = 1 if msc.n_words < 128 else spm                  # short site rows bump the sweep counter once per batch
        kernels_per_step = spm * msc.n_colours + bumps + 2 + 1   # sweeps + bumps, energy (2), label exchange (one fused launch)
    n_ladders = msc.n_ladders
    replicas_total = n_beta * n_ladders
    attempts_per_step = replicas_total * n * spm
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step()
        ms, t_wall0, t_wall1 = timed(step, args.steps)
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    value = attempts_per_step * args.steps / (ms * 1e-3)
    launches = args.steps * kernels_per_step

    # ---- dominant kernel alone: one colour of one sweep over this rank's slots -----------------------------------------
    n_time = max(4, spm)
    with torch.cuda.stream(stream):
        msc.sweep(2)
        sweep_ms_total, _, _ = timed(lambda: msc.sweep(n_time), 1)
    sweep_ms = sweep_ms_total / (n_time * msc.n_colours)             # average launch duration (incl. the 1-thread bumps)
    local_replicas = msc.n_beta * n_ladders
    attempts_per_launch = local_replicas * n / msc.n_colours
    peak, peak_src = measured_peaks()
    packed_bytes = 2.0 * (n / msc.n_colours) * msc.n_words * 4        # read the other colour once + write this colour
    traffic = ncu_traffic()
    kernel_name = (f"msc_sweep_kernel<5 steps + 4 merged, {'bit planes' if world > 1 else 'scalar thresholds'}> "
                   "(one colour of one sweep)")
    roofline = {"bound": "hbm", "kernel": kernel_name, "achieved": packed_bytes / (sweep_ms * 1e-3) / 1e9, "peak": peak,
                "unit": "GB/s", "frac": packed_bytes / (sweep_ms * 1e-3) / 1e9 / peak, "peak_source": peak_src,
                "traffic": (traffic or {}).get("dram_bytes_per_launch") if world == 1 else None,
                "traffic_source": (traffic or {}).get("source") if world == 1 else None,
                "launch_ms": sweep_ms, "attempts_per_launch": attempts_per_launch,
                "bytes_per_attempt": packed_bytes / attempts_per_launch,
                "share_of_step": (sweep_ms * spm * msc.n_colours) / (ms / args.steps),
                "note": "algorithmic bytes of the bit-packed layout (SURVEY 8d: 0.25 B/attempt = read the other colour's rows + "
                        "write this colour's, 1 bit per spin).  The kernel is bound by integer issue (Philox4x32-10 + bit-sliced "
                        "compare: the ALU pipe is the busiest unit), not by HBM: see roofline_issue.  SURVEY 8d's generic int8-CSR figure (42 B/attempt) does not "
                        "describe this layout; against it the same launch would read as "
                        f"{SURVEY_BYTES_PER_ATTEMPT * attempts_per_launch / (sweep_ms * 1e-3) / 1e9 / peak:.1f} x the HBM peak"}
    roofline_issue = None
    if world == 1 and traffic and traffic.get("warp_instructions_per_launch") and traffic.get("attempts_per_launch") == attempts_per_launch:
        props = torch.cuda.get_device_properties(device)
        sm_clock_hz = 1e6 * (clocks["sm_mhz"] if clocks and clocks.get("sm_mhz") else 1965.0)
        issue_peak = props.multi_processor_count * 4 * sm_clock_hz          # one warp instruction per scheduler and clock
        achieved = traffic["warp_instructions_per_launch"] / (sweep_ms * 1e-3)
        roofline_issue = {"bound": "issue", "achieved": achieved / 1e9, "peak": issue_peak / 1e9, "unit": "G warp-inst/s",
                          "frac": achieved / issue_peak,
                          "warp_instructions_per_attempt": traffic["warp_instructions_per_launch"] / attempts_per_launch,
                          "note": "SMs x 4 schedulers x SM clock under load; instruction count from the committed ncu capture "
                                  "of this kernel (profiles/traffic.json), duration measured live"}

    # ---- sustained: the same step for >= args.sustain seconds, clocks and power sampled under load ---------------------
    sustained = None
    if args.sustain > 0:
        s2 = ClockSampler(device) if rank == 0 else None
        chunk = max(1, int(0.5 / (ms / args.steps * 1e-3)))                  # about half a second per chunk
        n_chunks = max(1, int(np.ceil(args.sustain / (chunk * ms / args.steps * 1e-3))))
        with torch.cuda.stream(stream):
            sms, w0, w1 = timed(step, chunk * n_chunks)
        c2 = s2.stop(w0, w1) if s2 else None
        sustained = {"seconds": sms * 1e-3, "steps": chunk * n_chunks, "value": attempts_per_step * chunk * n_chunks / (sms * 1e-3),
                     "unit": UNIT, "clocks": c2}

    # ---- weak scaling beside the strong one (N > 1): every rank runs the WHOLE stated C5 on its own ladders -------------
    weak = None
    if world > 1:
        own = _lib.Msc(prob.inst, betas, args.n_ladders, seed=1000, ladder_offset=rank * n_ladders)
        own.set_stream(stream.cuda_stream)
        with torch.cuda.stream(stream):
            for _ in range(args.warmup):
                own.round(spm, pairs)
            wms, _, _ = timed(lambda: own.round(spm, pairs), args.steps)
        weak = {"value": attempts_per_step * world * args.steps / (wms * 1e-3), "unit": UNIT, "scaling": "weak",
                "replicas_total": replicas_total * world, "ms_per_step": wms / args.steps,
                "note": "independent ladders per GPU (N x the stated C5), no data-path collective"}
        own.close()

    # ---- end to end through the drop-in class: NPT(J, h, mode='production').run(...) ----------------------------------
    # host J (scipy CSR) in, numpy (M, Energy) out; one call = args.steps swap rounds.  Inside the timed region: CSR
    # upload, colouring, state initialisation, every round, the recorded last round and its float64 M on the host.
    def api_run(rounds):
        obj = NPT(A, np.zeros(n), mode="production", device=device)
        obj.num_runs = args.n_ladders
        obj.m_on_ranks = (0,)   # N > 1: the float64 M is materialised on rank 0 only (every rank gets Energy)
        return obj.run(betas, n_beta, [False] * n_beta, num_sweeps_MCMC=spm * rounds, num_sweeps_read=spm * rounds,
                       num_swap_attempts=rounds, num_swapping_pairs=pairs)

    api_run(max(1, args.warmup))
    barrier()
    t0 = time.perf_counter()
    M_out, E_out = api_run(args.steps)
    torch.cuda.synchronize(device)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    csr_bytes = A.indptr.nbytes + A.indices.nbytes + A.data.nbytes + 8 * n
    d2h = spm * n_beta * n + spm * n_beta * n_ladders * 8               # rank 0: recorded int8 states of run 0 (all slots) + energies
    e2e = {"value": attempts_per_step * args.steps / e2e_s, "unit": UNIT,
           "h2d_bytes_per_step": int(csr_bytes / args.steps), "d2h_bytes_per_step": int(d2h / args.steps),
           "ms_per_step": 1e3 * e2e_s / args.steps, "seconds_per_call": e2e_s, "rounds_per_call": args.steps,
           "api": "NPT(J, h, mode='production').run(beta_list, 32, [False]*32, ..., num_swap_attempts=steps) with num_runs ladders "
                  "side by side: host scipy J in, numpy M (float64, last round, run 0; on rank 0 when N > 1) and Energy out; state resident across rounds",
           "returned": {"M_shape": list(M_out.shape) if M_out is not None else None, "Energy_coldest": float(E_out[-1])},
           "numa_node_rank0": numa_node}
    del M_out

    # second key: one swap round per call through HOST buffers (the C-ABI entry point that ships the packed state both ways)
    e2e_host = None
    if world == 1 and not args.no_host_round:
        shape = msc.packed_shape()
        NH = max(1, args.e2e_handles)
        msc.set_stream(None)
        handles = [msc] + [_lib.Msc(prob.inst, betas, args.n_ladders, seed=8919 + i) for i in range(NH - 1)]
        h_in = [torch.empty(shape, dtype=torch.int32, pin_memory=True) for _ in range(NH)]
        h_out = [torch.empty(shape, dtype=torch.int32, pin_memory=True) for _ in range(NH)]
        h_E = [torch.empty((n_beta, n_ladders), dtype=torch.float64, pin_memory=True) for _ in range(NH)]
        msc.get_packed(h_in[0].numpy().view(np.uint32))
        for k in range(1, NH):
            h_in[k].copy_(h_in[0])

        def host_steps(count):
            for i in range(count):
                k = i % NH
                handles[k].sync()
                h_in[k], h_out[k] = h_out[k], h_in[k]
                handles[k].round_host_async(h_in[k].data_ptr(), spm, pairs, h_out[k].data_ptr(), h_E[k].data_ptr())
            for hdl in handles:
                hdl.sync()

        for k in range(NH):
            h_out[k].copy_(h_in[k])
        host_steps(max(NH, args.warmup))
        t0 = time.perf_counter()
        host_steps(args.steps)
        torch.cuda.synchronize(device)
        hs = time.perf_counter() - t0
        e2e_host = {"value": attempts_per_step * args.steps / hs, "unit": UNIT, "ms_per_step": 1e3 * hs / args.steps,
                    "h2d_bytes_per_step": int(h_in[0].numel() * 4),
                    "d2h_bytes_per_step": int(h_out[0].numel() * 4 + h_E[0].numel() * 8),
                    "api": f"nlmc_msc_round_host_async + nlmc_msc_sync (C ABI, pinned host buffers, {NH} handles in turn): the "
                           "whole packed state crosses PCIe both ways every round"}
        for hdl in handles[1:]:
            hdl.close()

    # ---- time to target (BASELINE metric, second half) and the CPU baseline: rank 0 at N = 1 -----------------------------
    ttt_rows, cpu = None, None
    if rank == 0 and world == 1 and not args.no_cpu:
        ttt = load_ttt()
        rows = ttt.measure(repeats=args.ttt_repeats, cpu=True)
        ttt_rows = []
        for name in ttt.ALL:
            g = [r for r in rows if r["instance"] == name and r["arm"].startswith("gpu")][0]
            c = [r for r in rows if r["instance"] == name and r["arm"].startswith("cpu")][0]
            ttt_rows.append({"instance": g["what"].replace("time to ground state, ", ""), "engine": g["engine"],
                             "gpu_median_s": g["median_seconds"], "cpu_median_s": c["median_seconds"],
                             "cpu_over_gpu": c["median_seconds"] / g["median_seconds"], "repeats": args.ttt_repeats,
                             "gpu_arm": g["arm"], "cpu_arm": c["arm"]})
        cores = os.cpu_count() or 1
        reps, sw = max(1, min(cores, 64)), max(1, args.ref_sweeps)
        once, attempts = cpu_port_rate(L, betas, reps, sw, cores)
        once()
        dt, used = once()
        cpu = {"value": attempts / dt, "unit": UNIT, "cores": used, "kind": "port",
               "sample": f"{reps} replicas x {sw} sweeps of the L={L} lattice ({attempts:.3g} attempts), "
                         "oracle/nlmc_oracle.c (reference algorithm, injected MT19937 stream)"}
    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
               "vs_baseline": None, "dtype": "u32", "dtype_note": "bit planes: 1 bit per spin, 32 ladders per 32-bit word",
               "data": "synthetic", "config": workload_config(args, world), "exchange": exchange,
               "roofline": roofline, "roofline_issue": roofline_issue, "cpu_baseline": cpu, "e2e": e2e,
               "e2e_host_buffers": e2e_host, "weak": weak, "sustained": sustained, "time_to_target": ttt_rows,
               "gpu_launches": launches, "clocks": clocks, "per_gpu_value": value / world,
               "ps_per_attempt": 1e12 / (value / world)}
        print(json.dumps(out), flush=True)
    if world == 1:
        msc.close()
    else:
        ens.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--L", type=int, default=64)
    ap.add_argument("--n-beta", dest="n_beta", type=int, default=32)
    ap.add_argument("--n-ladders", dest="n_ladders", type=int, default=128)
    ap.add_argument("--spm", type=int, default=16, help="sweeps per swap round (one step)")
    ap.add_argument("--pairs", type=int, default=10, help="swapping pairs per ladder per round (round(0.3*32), README)")
    ap.add_argument("--ref-sweeps", dest="ref_sweeps", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-handles", dest="e2e_handles", type=int, default=3,
                    help="handles used in turn by the host-buffer leg (copies of one batch overlap the sweeps of the others)")
    ap.add_argument("--no-host-round", dest="no_host_round", action="store_true", help="skip the host-buffer leg")
    ap.add_argument("--sustain", type=float, default=5.0, help="seconds of the sustained leg (0 = skip)")
    ap.add_argument("--ttt-repeats", dest="ttt_repeats", type=int, default=3)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world and world == 1 and args.gpus > 1:
        sys.exit("--gpus N > 1 must be launched with torchrun (one process per GPU), e.g.\n"
                 "  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 "
                 "--master-port 29500 bench.py --gpus N")
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
