"""Per-config measurements (BASELINE.json configs C1-C4; C5 is bench.py): spin-flip attempts/s of the CUDA
path in replay and production mode next to the CPU oracle port (the reference algorithm in C), on the GPU
box.  Prints one JSON line per measurement; the round's output is kept in profiles/.

    python tools/bench_configs.py [c1] [c2] [c3] [c4]
"""
import json
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "nonlocal-monte-carlo_b200"))
from nlmc_b200 import APT_ICM, NMC, NPT, APT_preprocessor, _lib, host, instances  # noqa: E402
from oracle import oracle as O  # noqa: E402  (CPU baseline legs only)

EPS = np.finfo(float).eps


def emit(**kw):
    print(json.dumps(kw), flush=True)


def cpu_mcmc_rate(csr, h, beta, sweeps=2):
    rs = np.random.RandomState(0)
    m0 = rs.choice([-1, 1], size=csr.n).astype(np.int8)
    perm, u = O.draw_sweeps(rs, sweeps, csr.n)
    t0 = time.perf_counter()
    O.mcmc(csr, h, m0, np.full(sweeps, beta), perm=perm, u=u)
    return sweeps * csr.n / (time.perf_counter() - t0)


def c1():
    """NMC.run, N=800 random +-1 graph (6% density), README parameters with sweeps cut to 1e3."""
    J, h = instances.random_pm_graph(800, 0.06, 1)
    args = (1000, 1000, 10, 1, 1, 20, 3, 3, 0.01, 0.9, 0.9999999, 0.999999, 100, EPS)
    attempts = (1000 + 30 * 1000) * 800
    for mode in ("replay", "production"):
        np.random.seed(1); random.seed(1)
        cwd = os.getcwd(); os.chdir("/tmp")
        t0 = time.perf_counter()
        M, E, mn = NMC(J, h, mode=mode).run(*args)
        dt = time.perf_counter() - t0
        os.chdir(cwd)
        emit(config="C1", what=f"NMC.run {mode}", seconds=dt, attempts=attempts, attempts_per_s=attempts / dt,
             min_energy=float(mn), M_shape=list(M.shape))
    csr = O.Csr(J)
    emit(config="C1", what="CPU oracle port MCMC, 1 core", attempts_per_s=cpu_mcmc_rate(csr, h, 3.0, 20),
         note="reference Python measured 8.6e3 attempts/s/core (BASELINE.md)")


def c2():
    """APT_preprocessor ladder + NPT on 3D +-J EA L=16, ~30 replicas."""
    A, h = instances.ea3d_pm_j(16, 2)
    n = 4096
    cwd = os.getcwd(); os.chdir("/tmp")
    np.random.seed(2); random.seed(2)
    t0 = time.perf_counter()
    beta, sigma = APT_preprocessor(A.copy(), h.copy(), mode="production").run(1000, 1000, 100, 0.5, 1.25, 1000, 64, 0, 8)
    dt = time.perf_counter() - t0
    emit(config="C2", what="APT_preprocessor.run production (README parameters)", seconds=dt, n_betas=len(beta),
         attempts=len(beta) * 100 * 1000 * n, attempts_per_s=len(beta) * 100 * 1000 * n / dt, beta_last=float(beta[-1]))
    betas = np.array(beta, dtype=float)[:30] if len(beta) >= 30 else np.linspace(0.5, 3.0, 30)
    R = len(betas)
    obj = NPT(A, h, mode="production")
    obj.num_runs = 128
    t0 = time.perf_counter()
    M, E = obj.run(betas, R, [False] * R, num_sweeps_MCMC=10000, num_sweeps_read=100, num_swap_attempts=10,
                   num_swapping_pairs=round(0.3 * R))
    dt = time.perf_counter() - t0
    att = 128 * R * n * 10000
    emit(config="C2", what="NPT.run production, README sweeps, 128 ladders in the bit lanes", seconds=dt, attempts=att,
         attempts_per_s=att / dt, best_energy=float(E.min()), best_energy_all_runs=float(obj.energies_all_runs.min()),
         note="includes building the reference's return value M: (30*4096) x 1000 float64 = 983 MB on the host")
    prob = host.Problem(A, h)
    msc = _lib.Msc(prob.inst, betas, 128, seed=1)
    msc.round(1000, round(0.3 * R)); msc.sync()
    t0 = time.perf_counter()
    for _ in range(5):
        msc.round(1000, round(0.3 * R))
    msc.sync()
    dt = (time.perf_counter() - t0) / 5
    emit(config="C2", what="one swap round (1000 sweeps + energies + exchange), device only", seconds=dt,
         attempts_per_s=128 * R * n * 1000 / dt)
    msc.close()
    # the config as BASELINE.json states it: doNMC on the 5 coldest replicas, README NMC parameters
    nmc_kw = dict(num_cycles=10, full_update_frequency=1, M_skip=1, temp_x=20, global_beta=1 / 0.366838 * 5, lambda_start=3,
                  lambda_end=0.01, lambda_reduction_factor=0.9, threshold_initial=0.9999999, threshold_cutoff=0.999999,
                  max_iterations=100, tolerance=EPS)
    doNMC = [False] * (R - 5) + [True] * 5
    for mode, sweeps in (("production", 10000), ("replay", 300)):
        np.random.seed(5); random.seed(5)
        t0 = time.perf_counter()
        M, E = NPT(A, h, mode=mode).run(betas, R, doNMC, num_sweeps_MCMC=sweeps, num_sweeps_read=100, num_swap_attempts=10,
                                        num_swapping_pairs=round(0.3 * R), **nmc_kw)
        dt = time.perf_counter() - t0
        phase = int(np.ceil(sweeps / 10 / 3 / 10))
        att = n * 10 * ((R - 5) * (sweeps // 10) + 5 * 30 * phase)
        emit(config="C2", what=f"NPT.run {mode}, doNMC on the 5 coldest, num_sweeps_MCMC={sweeps}", seconds=dt, attempts=att,
             attempts_per_s=att / dt, best_energy=float(E.min()))
    np.random.seed(3); random.seed(3)
    t0 = time.perf_counter()
    M, E = NPT(A, h, mode="replay").run(betas, R, [False] * R, num_sweeps_MCMC=200, num_sweeps_read=100,
                                        num_swap_attempts=10, num_swapping_pairs=round(0.3 * R))
    dt = time.perf_counter() - t0
    emit(config="C2", what="NPT.run replay (exact), sweeps cut to 200", seconds=dt, attempts=R * n * 200,
         attempts_per_s=R * n * 200 / dt)
    os.chdir(cwd)
    emit(config="C2", what="CPU oracle port MCMC, 1 core", attempts_per_s=cpu_mcmc_rate(O.Csr(A), h, 1.0, 10),
         note="reference Python measured 3.5e3-4.0e3 attempts/s/core (BASELINE.md)")


def c3():
    """SK N=2000 dense Gaussian, 64 betas x 32 runs = 2048 replicas on the tensor-core path."""
    J, h = instances.sk_gaussian(2000, 3)
    J = J / np.max(np.abs(J))
    prob = host.Problem(J, h)
    betas = np.tile(np.linspace(0.2, 3.0, 64), 32)
    for ns in (1, 3):
        d = _lib.Dense(prob.inst, betas, n_split=ns, seed=1)
        gemm_ms, sweep_ms = d.time_fields(20), d.time_sweeps(5)
        emit(config="C3", what=f"dense path, n_split={ns}", sweep_ms=sweep_ms, attempts_per_s=2048 * 2000 / sweep_ms * 1e3,
             full_field_gemm_ms=gemm_ms, gemm_tflops=2 * 2048 ** 3 * ns / gemm_ms / 1e9)
        d.close()
    emit(config="C3", what="CPU oracle port MCMC, 1 core", attempts_per_s=cpu_mcmc_rate(O.Csr(J), h, 1.0, 1),
         note="reference Python measured 1.5e2 attempts/s/core (BASELINE.md)")


def c4():
    """APT_ICM on 3D +-J EA L=32, reference semantics (10 sub-replicas per beta), reduced sweeps."""
    A, h = instances.ea3d_pm_j(32, 4)
    betas = np.linspace(0.3, 1.5, 8)
    np.random.seed(4); random.seed(4)
    cwd = os.getcwd(); os.chdir("/tmp")
    t0 = time.perf_counter()
    M, E = APT_ICM(A, h, mode="replay").run(betas, 8, num_sweeps_MCMC=4, num_sweeps_read=2, num_swap_attempts=2,
                                            num_swapping_pairs=2)
    dt = time.perf_counter() - t0
    os.chdir(cwd)
    att = 8 * 10 * 32768 * 4
    emit(config="C4", what="APT_ICM.run replay: 8 betas x 10 sub-replicas, 2 rounds x 2 sweeps, 80 Houdayer pairs",
         seconds=dt, attempts=att, attempts_per_s=att / dt, energies=[float(e) for e in E])
    betas32 = np.linspace(0.2, 1.6, 32)
    np.random.seed(4); random.seed(4)
    os.chdir("/tmp")
    for warm in (True, False):  # first call creates handles/graphs
        t0 = time.perf_counter()
        M, E = APT_ICM(A, h, mode="production").run(betas32, 32, num_sweeps_MCMC=1000, num_sweeps_read=1000,
                                                   num_swap_attempts=100, num_swapping_pairs=10)
        dt = time.perf_counter() - t0
    os.chdir(cwd)
    att = 32 * 10 * 32768 * 1000
    emit(config="C4", what="APT_ICM.run production: 32 betas x 10 sub-replicas, 100 rounds x 10 sweeps, device-resident rounds",
         seconds=dt, attempts=att, attempts_per_s=att / dt, min_energy=float(E.min()), M_shape=list(M.shape))
    rs = np.random.RandomState(1)
    s1 = rs.choice([-1, 1], size=(40, 32768)).astype(np.int8)
    s2 = np.where(rs.rand(40, 32768) < 0.3, -s1, s1).astype(np.int8)
    prob = host.Problem(A, h)
    _lib.icm_clusters(prob.inst, s1[:2], s2[:2])
    t0 = time.perf_counter()
    labels, counts = _lib.icm_clusters(prob.inst, s1, s2)
    dt = time.perf_counter() - t0
    emit(config="C4", what="K7: disagreement clusters of 40 pairs at L=32 (one launch, incl. copies)", seconds=dt,
         seconds_per_pair=dt / 40, clusters_mean=float(counts.mean()),
         note="reference Python measured 0.19 s per pair at L=12 (N=1728)")
    t0 = time.perf_counter()
    O.disagreement_clusters(O.Csr(A), s1[0], s2[0])
    emit(config="C4", what="CPU oracle port, one pair at L=32", seconds_per_pair=time.perf_counter() - t0)


if __name__ == "__main__":
    host.Problem(np.array([[0.0, 1.0], [1.0, 0.0]]), np.zeros(2))  # create the CUDA context once, outside every timing
    which = [a.lower() for a in sys.argv[1:]] or ["c1", "c2", "c3", "c4"]
    for name in which:
        {"c1": c1, "c2": c2, "c3": c3, "c4": c4}[name]()
