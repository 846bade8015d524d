#!/usr/bin/env python
"""bench.py -- headline benchmark of nlmc_b200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Metric (BASELINE.json): spin-flip attempts/s.  Workload: config C5 -- 3D +-J Edwards-Anderson L=64
(262,144 spins), NPT with 32 inverse temperatures x 128 independent ladders = 4096 replicas per GPU.
A "step" is one swap round of NPT.run for all ladders: `spm` heat-bath sweeps of every replica, the
energies of every replica, and the replica-exchange step (NPT/npt.py:617-680).

  value   attempts/s with the state resident in HBM, CUDA events on the library's stream, max over ranks.
  e2e     the same step through the C-ABI entry point that takes HOST buffers (nlmc_msc_round_host):
          packed spins in from pinned host memory, packed spins + energies back out, copies inside the
          timed region (the reference ships m_start to its workers and M back every round, npt.py:625-644).
  N > 1   one process per GPU (torchrun); ladders are independent, so every rank runs its own 128 ladders
          (weak scaling, no data-path collective); only the timing is reduced (max) over NCCL.
  --impl reference   the reference algorithm on the host cores: the oracle C port of MCMC
          (oracle/nlmc_oracle.c; the reference itself is pure Python and cannot travel to the GPU box),
          all host threads, a bounded sample of the same workload per step.
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
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": workload_config(args, world, reference=True),
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def workload_config(args, world, reference=False):
    return {"workload": f"C5: 3D +-J Edwards-Anderson L={args.L} ({args.L ** 3} spins) NPT, {args.n_beta} betas x "
                        f"{args.n_ladders} ladders = {args.n_beta * args.n_ladders} replicas per GPU, "
                        f"{args.spm} sweeps per swap round, {args.pairs} swapping pairs per ladder",
            "L": args.L, "n_beta": args.n_beta, "n_ladders_per_gpu": args.n_ladders, "sweeps_per_step": args.spm,
            "replicas_total": args.n_beta * args.n_ladders * world, "beta_range": [0.2, 2.0],
            "parallelism": f"replicas: {world} x {args.n_ladders} independent ladders, no data-path collective",
            "l2": "state (n x words x 4 B = 134 MB at the default size) is larger than the 126 MB L2; no flush needed "
                  "(ncu: 78 MB of DRAM reads per colour launch against 67 MB compulsory, profiles/r1c_sweep_kernel_summary.md)"
            if not reference else "n/a (host)"}


def bind_to_gpu_numa_node(device_index: int):
    """Pin this rank to the CPUs of its GPU's NUMA node so that the pinned host buffers of the end-to-end leg are
    first-touched next to the GPU's PCIe root (matters when 8 ranks stream 268 MB per step each)."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
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


def run_ours(args, rank, world, local_rank):
    import torch
    from nlmc_b200 import _lib, host
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    device = local_rank
    _lib.require_device(device)
    numa_node = bind_to_gpu_numa_node(device) if world > 1 else None
    L, n_beta, n_ladders, spm, pairs = args.L, args.n_beta, args.n_ladders, args.spm, args.pairs
    betas = np.linspace(0.2, 2.0, n_beta)
    sampler = ClockSampler(device) if rank == 0 else None  # started early: nvidia-smi needs ~0.1 s before its first row
    A = ea3d_csr(L, 5)
    prob = host.Problem(A, np.zeros(A.shape[0]), device=device)
    # every rank owns its own block of ladders; streams are keyed by the global ladder index
    msc = _lib.Msc(prob.inst, betas, n_ladders, seed=1000, ladder_offset=rank * (((n_ladders + 127) // 128) * 128))
    n = prob.n
    replicas = n_beta * msc.n_ladders
    attempts_per_step = replicas * n * spm

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(device)

    # ---- device-resident throughput --------------------------------------------------------------
    for _ in range(args.warmup):
        msc.round(spm, pairs)
    msc.sync()
    barrier()
    t_wall0 = time.perf_counter()
    msc.timer_mark(0)
    for _ in range(args.steps):
        msc.round(spm, pairs)
    msc.timer_mark(1)
    msc.sync()
    ms = msc.timer_elapsed_ms()
    t_wall1 = time.perf_counter()
    barrier()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    if dist is not None:
        t = torch.tensor([ms], device=f"cuda:{device}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = attempts_per_step * world * args.steps / (ms * 1e-3)
    # per step: spm x (colour kernels + sweep-counter bump) + energy (2) + exchange (2) + round-counter bump,
    # replayed from one captured CUDA graph
    launches = args.steps * (spm * (msc.n_colours + 1) + 5)

    # ---- dominant kernel alone: the colour sweep ---------------------------------------------------
    n_time = max(4, spm)
    msc.sweep(2)
    msc.sync()
    msc.timer_mark(0)
    msc.sweep(n_time)
    msc.timer_mark(1)
    msc.sync()
    sweep_ms = msc.timer_elapsed_ms() / (n_time * msc.n_colours)  # average launch duration
    attempts_per_launch = replicas * n / msc.n_colours
    peak, peak_src = measured_peaks()
    survey_bytes = SURVEY_BYTES_PER_ATTEMPT * attempts_per_launch
    packed_bytes = 2.0 * (n / msc.n_colours) * msc.n_words * 4  # read the other colour once + write this colour
    traffic = ncu_traffic()
    roofline = {"bound": "hbm", "kernel": "msc_sweep_kernel (one colour of one sweep)",
                "achieved": survey_bytes / (sweep_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                "frac": survey_bytes / (sweep_ms * 1e-3) / 1e9 / peak, "peak_source": peak_src,
                "traffic": None if traffic is None else traffic.get("dram_bytes_per_launch"),
                "launch_ms": sweep_ms, "attempts_per_launch": attempts_per_launch,
                "bytes_per_attempt": SURVEY_BYTES_PER_ATTEMPT,
                "note": "achieved uses SURVEY.md 8(d)'s generic-CSR figure (42 B/attempt); the kernel is bit-packed "
                        "(32 ladders per word), so its own traffic is far lower -- see roofline_packed and DESIGN.md",
                "share_of_step": (sweep_ms * spm * msc.n_colours) / (ms / args.steps)}
    roofline_packed = {"bound": "hbm", "achieved": packed_bytes / (sweep_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                       "frac": packed_bytes / (sweep_ms * 1e-3) / 1e9 / peak,
                       "bytes_per_attempt": packed_bytes / attempts_per_launch,
                       "note": "bit-packed layout: 0.25 B/attempt compulsory traffic; the kernel is ALU/issue bound "
                               "(Philox + bit-sliced logic), not HBM bound"}

    # what actually bounds the kernel: warp-instruction issue (integer pipes).  The instruction count per launch is a
    # property of the code (ncu smsp__inst_executed.sum of the committed capture); the launch duration is measured here.
    roofline_issue = None
    if traffic and traffic.get("warp_instructions_per_launch") and traffic.get("attempts_per_launch") == attempts_per_launch:
        props = torch.cuda.get_device_properties(device)
        sm_clock_hz = 1e6 * (clocks["sm_mhz"] if clocks and clocks.get("sm_mhz") else 1965.0)
        issue_peak = props.multi_processor_count * 4 * sm_clock_hz          # one warp instruction per scheduler and clock
        achieved = traffic["warp_instructions_per_launch"] / (sweep_ms * 1e-3)
        roofline_issue = {"bound": "issue", "achieved": achieved / 1e9, "peak": issue_peak / 1e9, "unit": "G warp-inst/s",
                          "frac": achieved / issue_peak,
                          "warp_instructions_per_attempt": traffic["warp_instructions_per_launch"] / attempts_per_launch,
                          "note": "SMs x 4 schedulers x SM clock under load; instruction count from the committed ncu capture "
                                  "(profiles/traffic.json), duration measured live"}

    # ---- end to end through the host-buffer entry point -------------------------------------------
    # Every step ships one batch of packed states from pinned host memory to the device, runs the round and
    # ships states + energies back.  NH handles are used in turn (nlmc_msc_round_host_async + nlmc_msc_sync), so
    # the copies of one batch overlap the sweeps of the others; each handle works strictly in -> round -> out.
    # Measured with tools/e2e_pipeline_probe.py: 11.1 / 7.6 / 6.2 / 6.1 ms per step with 1 / 2 / 3 / 4 handles.
    shape = msc.packed_shape()
    NH = max(1, args.e2e_handles)
    handles = [msc] + [_lib.Msc(prob.inst, betas, n_ladders, seed=8919 + i,
                                ladder_offset=rank * (((n_ladders + 127) // 128) * 128)) for i in range(NH - 1)]
    h_in = [torch.empty(shape, dtype=torch.int32, pin_memory=True) for _ in range(NH)]
    h_out = [torch.empty(shape, dtype=torch.int32, pin_memory=True) for _ in range(NH)]
    h_E = [torch.empty((n_beta, msc.n_ladders), dtype=torch.float64, pin_memory=True) for _ in range(NH)]
    msc.get_packed(h_in[0].numpy().view(np.uint32))  # current device state -> pinned host buffers
    for k in range(1, NH):
        h_in[k].copy_(h_in[0])

    def e2e_steps(count):
        for i in range(count):
            k = i % NH
            handles[k].sync()                      # batch i-NH is complete: its outputs are on the host
            h_in[k], h_out[k] = h_out[k], h_in[k]  # and become the next input of this handle
            handles[k].round_host_async(h_in[k].data_ptr(), spm, pairs, h_out[k].data_ptr(), h_E[k].data_ptr())
        for hdl in handles:
            hdl.sync()

    for k in range(NH):
        h_out[k].copy_(h_in[k])
    e2e_steps(max(NH, args.warmup))
    barrier()
    t0 = time.perf_counter()
    e2e_steps(args.steps)
    torch.cuda.synchronize(device)
    e2e_s = time.perf_counter() - t0
    barrier()
    if dist is not None:
        t = torch.tensor([e2e_s], device=f"cuda:{device}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    # one batch alone, no overlap: the latency of a single host -> host round
    handles[0].sync()
    t1 = time.perf_counter()
    handles[0].round_host(h_in[0].data_ptr(), spm, pairs, h_out[0].data_ptr(), h_E[0].data_ptr())
    single_ms = 1e3 * (time.perf_counter() - t1)
    e2e = {"value": attempts_per_step * world * args.steps / e2e_s, "unit": UNIT,
           "h2d_bytes_per_step": int(h_in[0].numel() * 4),
           "d2h_bytes_per_step": int(h_out[0].numel() * 4 + h_E[0].numel() * 8),
           "ms_per_step": 1e3 * e2e_s / args.steps,
           "api": f"nlmc_msc_round_host_async + nlmc_msc_sync (C ABI, pinned host buffers, {NH} handles in turn)",
           "single_batch_ms": single_ms,
           "numa_node_rank0": numa_node,
           "mean_energy_coldest": float(h_E[0][-1].mean())}
    for hdl in handles[1:]:
        hdl.close()

    # ---- CPU baseline (rank 0, N = 1 only): the oracle port on a bounded sample --------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
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
               "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "u32", "dtype_note": "bit planes: 1 bit per spin, 32 ladders per 32-bit word", "data": "synthetic",
               "config": workload_config(args, world), "roofline": roofline, "roofline_packed": roofline_packed, "roofline_issue": roofline_issue,
               "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
               "per_gpu_value": value / world, "ps_per_attempt": 1e12 / (value / world)}
        print(json.dumps(out), flush=True)
    msc.close()
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
                    help="handles used in turn by the end-to-end leg (copies of one batch overlap the sweeps of the others)")
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
