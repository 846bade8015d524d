import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "nonlocal-monte-carlo_b200"))
import numpy as np
from nlmc_b200 import _lib, host
from nlmc_b200 import instances as O  # generators of the benchmark instances
which = sys.argv[1] if len(sys.argv) > 1 else "ea"
J, h = O.ea3d_pm_j(16, 2) if which == "ea" else O.random_pm_graph(800, 0.06, 1)
beta = 1 / 0.366838 * 5 if which == "ea" else 3.0
prob = host.Problem(J, h)
lbp = _lib.Lbp(prob.inst)
ms = np.random.RandomState(0).choice([-1.0, 1.0], size=prob.n)
lbp.reset(ms)
for lam in (3.0, 2.7, 2.43):
    print(lbp.step(lam, beta, np.finfo(float).eps, 100)[1])
