import os, sys
sys.path.insert(0, "nonlocal-monte-carlo_b200")
import numpy as np
from nlmc_b200 import _lib, host, instances
os.environ["NLMC_DENSE_FUSED_PROF"] = "1"
J, h = instances.sk_gaussian(2000, 3); J = J / np.max(np.abs(J))
prob = host.Problem(J, h)
d = _lib.Dense(prob.inst, np.tile(np.linspace(0.2, 3.0, 64), 32), n_split=3, seed=1)
d.sweep(5); d.sync()
print("ms/sweep", d.time_sweeps(10))
d.close() if hasattr(d, "close") else None
del d
