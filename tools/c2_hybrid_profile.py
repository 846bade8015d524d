"""cProfile of config C2 as stated (NPT.run production with doNMC on the 5 coldest replicas, fixed ladder): where the time goes.
    python tools/c2_hybrid_profile.py"""
import cProfile
import os
import pstats
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nonlocal-monte-carlo_b200"))
from nlmc_b200 import NPT, instances  # noqa: E402

EPS = np.finfo(float).eps
A, h = instances.ea3d_pm_j(16, 2)
betas = np.linspace(0.5, 3.0, 30)
R = 30
nmc_kw = dict(num_cycles=10, full_update_frequency=1, M_skip=1, temp_x=20, global_beta=1 / 0.366838 * 5, lambda_start=3,
              lambda_end=0.01, lambda_reduction_factor=0.9, threshold_initial=0.9999999, threshold_cutoff=0.999999,
              max_iterations=100, tolerance=EPS)
doNMC = [False] * (R - 5) + [True] * 5
os.chdir("/tmp")


def call():
    np.random.seed(5); random.seed(5)
    return NPT(A, h, mode="production").run(betas, R, doNMC, num_sweeps_MCMC=10000, num_sweeps_read=100, num_swap_attempts=10,
                                            num_swapping_pairs=round(0.3 * R), **nmc_kw)


call()
t0 = time.perf_counter()
call()
print(f"wall {time.perf_counter() - t0:.3f} s")
pr = cProfile.Profile()
pr.enable()
call()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
