import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "nonlocal-monte-carlo_b200"))
import numpy as np
from nlmc_b200 import _lib, host, nmc_core
from nlmc_b200 import instances as O  # generators of the benchmark instances
eps = np.finfo(float).eps
for name, (J, h), beta in (("EA L=16", O.ea3d_pm_j(16, 2), 1 / 0.366838 * 5), ("C1 N=800", O.random_pm_graph(800, 0.06, 1), 3.0)):
    prob = host.Problem(J, h)
    lbp = _lib.Lbp(prob.inst)
    rs = np.random.RandomState(0)
    ms = rs.choice([-1.0, 1.0], size=prob.n)
    trace = []
    t0 = time.perf_counter()
    cl = nmc_core.lbp_convexified(prob, lbp, ms, 3, 0.01, 0.9, eps, 100, 0.9999999, 0.999999, beta, trace=trace)
    dt = time.perf_counter() - t0
    iters = sum(t[1] + 1 for t in trace)
    print(f"{name}: lbp_convexified {dt*1e3:.1f} ms, {len(trace)} lambda steps, {iters} iterations, {dt/iters*1e6:.1f} us/iteration, backbone {sum(len(c) for c in cl)}")
