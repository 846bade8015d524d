import cProfile, os, pstats, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "nonlocal-monte-carlo_b200"))
import numpy as np
from nlmc_b200 import NPT, _lib, host
from nlmc_b200 import instances as O  # generators of the benchmark instances
A, h = O.ea3d_pm_j(16, 2)
betas = np.linspace(0.5, 3.0, 30)
os.chdir("/tmp")
obj = NPT(A, h, mode="production"); obj.num_runs = 128
pr = cProfile.Profile(); pr.enable()
M, E = obj.run(betas, 30, [False] * 30, num_sweeps_MCMC=10000, num_sweeps_read=100, num_swap_attempts=10, num_swapping_pairs=9)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(12)
prob = host.Problem(A, h)
msc = _lib.Msc(prob.inst, betas, 128, seed=1)
for n in (1000, 1000, 1000):
    msc.sync(); t0 = time.perf_counter(); msc.round(n, 9); msc.sync(); print("round(1000) s:", time.perf_counter() - t0)
msc.sync(); t0 = time.perf_counter(); msc.sweep_record(1000, 0, True); print("sweep_record(1000) s:", time.perf_counter() - t0)
