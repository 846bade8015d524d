"""cProfile of config C5 through NPT(J, h, mode='production').run (the e2e leg of bench.py): where the host time goes.
    python tools/c5_api_profile.py [rounds]"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nonlocal-monte-carlo_b200"))
from nlmc_b200 import NPT, instances  # noqa: E402

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 20
A, h = instances.ea3d_pm_j(64, 5)
betas = np.linspace(0.2, 2.0, 32)


def call(k):
    obj = NPT(A, h, mode="production")
    obj.num_runs = 128
    return obj.run(betas, 32, [False] * 32, num_sweeps_MCMC=16 * k, num_sweeps_read=16 * k, num_swap_attempts=k,
                   num_swapping_pairs=10)


call(3)
t0 = time.perf_counter()
call(rounds)
print(f"wall {time.perf_counter() - t0:.3f} s for {rounds} rounds")
pr = cProfile.Profile()
pr.enable()
call(rounds)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
