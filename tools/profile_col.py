"""K2a (graph-coloured sparse sweep) micro-benchmark on the C1-shaped graph (n=800, density 0.06, +-1 couplings).
`python tools/profile_col.py` prints us/sweep for few and many replicas; `--ncu` runs the short fixed sequence that the
ncu capture in profiles/r1_other_kernels_summary.md was taken on."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "nonlocal-monte-carlo_b200"))
import numpy as np
from nlmc_b200 import _lib, host
from nlmc_b200 import instances as O  # generators of the benchmark instances

J, h = O.random_pm_graph(800, 0.06, 1)
prob = host.Problem(J, h)
if "--ncu" in sys.argv:
    c = _lib.Col(prob.inst, np.linspace(0.2, 3.0, 1184), seed=1)
    c.sweep(20); c.sync(); c.sweep(100); c.sync()
    print("ok", c.n_colours)
    sys.exit(0)
for R in (1, 148, 296, 592, 1184, 4736):
    c = _lib.Col(prob.inst, np.linspace(0.2, 3.0, R), seed=1)
    c.sweep(200); c.sync()
    t0 = time.perf_counter(); c.sweep(2000); c.sync(); dt = time.perf_counter() - t0
    print(f"R={R}: {dt / 2000 * 1e6:.2f} us/sweep, {R * 800 * 2000 / dt:.3e} attempts/s, "
          f"colours={c.n_colours} smem_csr={c.csr_in_smem}")
