"""Dev tool: config C5 (3D +-J EA L=64, 32 betas x 128 ladders) through the drop-in NPT class in production mode."""
import os, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "nonlocal-monte-carlo_b200"))
import numpy as np
from nlmc_b200 import NPT
from bench import ea3d_csr
os.chdir(tempfile.mkdtemp())
A = ea3d_csr(64, 5)
betas = np.linspace(0.2, 2.0, 32)
np.random.seed(1)
obj = NPT(A, np.zeros(A.shape[0]), mode="production")
obj.num_runs = 128
for sweeps in (160, 1600):
    t0 = time.perf_counter()
    M, E = obj.run(betas, 32, [False] * 32, num_sweeps_MCMC=sweeps, num_sweeps_read=sweeps, num_swap_attempts=sweeps // 16,
                   num_swapping_pairs=10)
    dt = time.perf_counter() - t0
    att = 4096 * 262144 * sweeps
    print(f"C5 NPT.run production, {sweeps} sweeps: {dt:.2f} s, {att / dt:.3e} attempts/s, M {M.shape}, "
          f"E/N coldest {E[-1] / 262144:.4f}, all-runs energies {obj.energies_all_runs.shape}", flush=True)
