"""Dev tool: C3-sized dense run (2048 replicas x N=2000) for ncu launch lists / captures."""
import sys
sys.path.insert(0, "nonlocal-monte-carlo_b200"); sys.path.insert(0, ".")
import numpy as np
from nlmc_b200 import _lib, host
from nlmc_b200 import instances as O  # generators of the benchmark instances
n_split = int(sys.argv[1]) if len(sys.argv) > 1 else 3
J, h = O.sk_gaussian(2000, 3); J = J / np.max(np.abs(J))
prob = host.Problem(J, h)
d = _lib.Dense(prob.inst, np.tile(np.linspace(0.2, 3.0, 64), 32), n_split=n_split, seed=1)
d.sweep(2); d.fields(fetch=False); d.fields(fetch=False); d.sync()
print("sweep ms", d.time_sweeps(2), "gemm ms", d.time_fields(10))
