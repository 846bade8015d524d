import os, sys
sys.path.insert(0, "nonlocal-monte-carlo_b200")
import numpy as np
from nlmc_b200 import _lib, host, instances
J, h = instances.sk_gaussian(2000, 3); J = J / np.max(np.abs(J))
prob = host.Problem(J, h)
d = _lib.Dense(prob.inst, np.tile(np.linspace(0.2, 3.0, 64), 32), n_split=3, seed=1)
d.sweep(3); d.sync()
d.close()
