"""Dev tool: cProfile of NMC.run (C1, replay) to see where the wall time goes."""
import cProfile, os, pstats, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "nonlocal-monte-carlo_b200"))
import numpy as np
from nlmc_b200 import NMC
from nlmc_b200 import instances as O  # generators of the benchmark instances
J, h = O.random_pm_graph(800, 0.06, 1)
mode = sys.argv[1] if len(sys.argv) > 1 else "replay"
args = (1000, 1000, 10, 1, 1, 20, 3, 3, 0.01, 0.9, 0.9999999, 0.999999, 100, np.finfo(float).eps)
np.random.seed(1); random.seed(1)
os.chdir("/tmp")
pr = cProfile.Profile(); pr.enable()
NMC(J, h, mode=mode).run(*args)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
