"""Dev tool: C1 NMC.run in production mode (reduced sweeps) for ncu launch lists."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "nonlocal-monte-carlo_b200"))
import numpy as np
from nlmc_b200 import NMC
from nlmc_b200 import instances as O  # generators of the benchmark instances
os.chdir(tempfile.mkdtemp())
J, h = O.random_pm_graph(800, 0.06, 1)
np.random.seed(1)
M, E, mn = NMC(J, h, mode="production").run(1000, 1000, 3, 1, 1, 20, 3, 3, 0.01, 0.9, 0.9999999, 0.999999, 100, np.finfo(float).eps)
print("min energy", mn, M.shape)
