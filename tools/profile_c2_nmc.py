import cProfile, os, pstats, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "nonlocal-monte-carlo_b200"))
import numpy as np
from nlmc_b200 import NPT, host
from nlmc_b200 import instances as O  # generators of the benchmark instances
A, h = O.ea3d_pm_j(16, 2)
betas = np.linspace(0.5, 3.0, 30)
host.Problem(np.array([[0.0, 1.0], [1.0, 0.0]]), np.zeros(2))
os.chdir("/tmp"); np.random.seed(5); random.seed(5)
kw = dict(num_cycles=10, full_update_frequency=1, M_skip=1, temp_x=20, global_beta=1 / 0.366838 * 5, lambda_start=3, lambda_end=0.01,
          lambda_reduction_factor=0.9, threshold_initial=0.9999999, threshold_cutoff=0.999999, max_iterations=100, tolerance=np.finfo(float).eps)
import time; t0 = time.perf_counter(); NPT(A, h, mode="production").run(betas, 30, [False] * 25 + [True] * 5, num_sweeps_MCMC=10000, num_sweeps_read=100, num_swap_attempts=10, num_swapping_pairs=9, **kw); print("first call", time.perf_counter() - t0); t0 = time.perf_counter(); NPT(A, h, mode="production").run(betas, 30, [False] * 25 + [True] * 5, num_sweeps_MCMC=10000, num_sweeps_read=100, num_swap_attempts=10, num_swapping_pairs=9, **kw); print("second call", time.perf_counter() - t0)
pr = cProfile.Profile(); pr.enable()
NPT(A, h, mode="production").run(betas, 30, [False] * 25 + [True] * 5, num_sweeps_MCMC=10000, num_sweeps_read=100, num_swap_attempts=10, num_swapping_pairs=9, **kw)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
