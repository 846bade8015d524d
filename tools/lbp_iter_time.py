import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "nonlocal-monte-carlo_b200"))
import numpy as np
from nlmc_b200 import _lib, host
from nlmc_b200 import instances as O  # generators of the benchmark instances
for name, (J, h), beta in (("EA L=16", O.ea3d_pm_j(16, 2), 13.6), ("C1 N=800", O.random_pm_graph(800, 0.06, 1), 3.0), ("EA L=32", O.ea3d_pm_j(32, 4), 3.0)):
    prob = host.Problem(J, h)
    lbp = _lib.Lbp(prob.inst)
    ms = np.random.RandomState(0).choice([-1.0, 1.0], size=prob.n)
    lbp.reset(ms)
    lbp.step(3.0, beta, -1.0, 2)
    res = []
    for iters in (1, 11, 101):
        lbp.reset(ms)
        t0 = time.perf_counter(); lbp.step(3.0, beta, -1.0, iters); res.append(time.perf_counter() - t0)
    print(f"{name}: launch+1 iter {res[0]*1e6:.0f} us; per iteration {(res[2]-res[1])/90*1e6:.1f} us (n={prob.n}, nnz={len(prob.val)})", flush=True)
