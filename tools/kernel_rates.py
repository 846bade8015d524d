"""Dev tool: measured rates of the smaller kernels against their algorithmic bytes.
K4 energy_kernel (states/s, GB/s on nnz*13 B), K5 lbp_kernel (us per iteration, GB/s on 32*nnz B), K7 icm (pairs/s)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "nonlocal-monte-carlo_b200"))
import json
import numpy as np
from nlmc_b200 import _lib, host, nmc_core
from nlmc_b200 import instances as O  # generators of the benchmark instances
def emit(**kw): print(json.dumps(kw), flush=True)
eps = np.finfo(float).eps
host.Problem(np.array([[0.0, 1.0], [1.0, 0.0]]), np.zeros(2))  # CUDA context
# K4: energies of many states
for name, (J, h) in (("C1 N=800 deg 48", O.random_pm_graph(800, 0.06, 1)), ("EA L=32", O.ea3d_pm_j(32, 4))):
    prob = host.Problem(J, h)
    rs = np.random.RandomState(0)
    S = rs.choice([-1, 1], size=(4096, prob.n)).astype(np.int8)
    prob.inst.energy_states(S[:8])
    t0 = time.perf_counter(); prob.inst.energy_states(S); dt = time.perf_counter() - t0
    nnz = len(prob.val)
    emit(kernel="K4 energy_kernel (incl. H2D of the states)", instance=name, states=len(S), seconds=dt, states_per_s=len(S) / dt,
         algorithmic_GBps=len(S) * (nnz * 13 + prob.n) / dt / 1e9)
# K5: LBP iterations
for name, (J, h), beta in (("EA L=16", O.ea3d_pm_j(16, 2), 1 / 0.366838 * 5), ("C1 N=800", O.random_pm_graph(800, 0.06, 1), 3.0)):
    prob = host.Problem(J, h)
    lbp = _lib.Lbp(prob.inst)
    ms = np.random.RandomState(0).choice([-1.0, 1.0], size=prob.n)
    trace = []
    nmc_core.lbp_convexified(prob, lbp, ms, 3, 0.01, 0.9, eps, 100, 0.9999999, 0.999999, beta)
    t0 = time.perf_counter()
    nmc_core.lbp_convexified(prob, lbp, ms, 3, 0.01, 0.9, eps, 100, 0.9999999, 0.999999, beta, trace=trace)
    dt = time.perf_counter() - t0
    iters = sum(t[1] + 1 for t in trace)
    nnz = len(prob.val)
    emit(kernel="K5 lbp_kernel", instance=name, lambda_steps=len(trace), iterations=iters, seconds=dt, us_per_iteration=dt / iters * 1e6,
         algorithmic_GBps=iters * 32 * nnz / dt / 1e9, note="2 grid-wide syncs per iteration: latency-bound on these sizes")
# K7: Houdayer clusters
A, h = O.ea3d_pm_j(32, 4)
prob = host.Problem(A, h)
rs = np.random.RandomState(1)
for P in (40, 640):
    s1 = rs.choice([-1, 1], size=(P, prob.n)).astype(np.int8)
    s2 = np.where(rs.rand(P, prob.n) < 0.3, -s1, s1).astype(np.int8)
    _lib.icm_clusters(prob.inst, s1[:2], s2[:2])
    t0 = time.perf_counter(); _lib.icm_clusters(prob.inst, s1, s2); dt = time.perf_counter() - t0
    emit(kernel="K7 icm_components_kernel (incl. copies)", instance="EA L=32", pairs=P, seconds=dt, pairs_per_s=P / dt)
