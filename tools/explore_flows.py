"""Dev tool: run the four classes in both modes on the instance kinds of the reference's example scripts and report
exceptions (used to find unsupported combinations)."""
import os, sys, tempfile, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "nonlocal-monte-carlo_b200"))
import numpy as np
import scipy.sparse as sp
from nlmc_b200 import NMC, NPT, APT_preprocessor, APT_ICM
from nlmc_b200 import instances as O  # generators of the benchmark instances
os.chdir(tempfile.mkdtemp())
eps = np.finfo(float).eps

def gauss(N, column_h=True):
    h = np.random.randn(N, 1) if column_h else np.random.randn(N)
    iu = np.triu_indices(N, 1); J = np.zeros((N, N)); J[iu] = np.random.randn(len(iu[0])); J += J.T
    return J, h

cases = {}
np.random.seed(1)
J, h = gauss(10); cases["gauss10_dense_colh"] = (J, h)
cases["gauss10_csr_colh"] = (sp.csr_matrix(J), h)
J2, h2 = gauss(40, column_h=False); cases["gauss40_dense_flat_h"] = (J2, h2)
A, hz = O.ea3d_pm_j(4, 3); cases["ea_L4_csr"] = (A, hz)
cases["ea_L4_dense"] = (A.toarray(), hz)
Jg, hg = O.random_pm_graph(60, 0.15, 2); cases["pm_graph60"] = (Jg, hg)
Jf = Jg.copy(); cases["pm_graph60_field"] = (Jf, 0.3 * np.random.randn(60))

for name, (J, h) in cases.items():
    for mode in ("replay", "production"):
        for what in ("prep", "npt", "npt_nmc", "nmc", "icm"):
            np.random.seed(3); random.seed(3)
            try:
                Jc = J.copy(); hc = np.array(h, dtype=float).copy()
                if what == "prep":
                    b, s = APT_preprocessor(Jc, hc, mode=mode).run(num_sweeps_MCMC=20, num_sweeps_read=20, num_rng=4, beta_start=0.5,
                                                                   alpha=1.25, sigma_E_val=1000, beta_max=3, use_hash_table=0, num_cores=1)
                    assert len(b) >= 1
                elif what in ("npt", "npt_nmc"):
                    betas = np.linspace(0.3, 2.0, 5)
                    Jn = Jc.toarray() if (what == "npt_nmc" and sp.issparse(Jc) and mode == "never") else Jc
                    M, E = NPT(Jn, hc, mode=mode).run(betas, 5, [False] * 3 + [what == "npt_nmc"] * 2, num_sweeps_MCMC=40, num_sweeps_read=20,
                                                      num_swap_attempts=4, num_swapping_pairs=2, num_cycles=2, global_beta=2.0, lambda_start=3,
                                                      threshold_initial=0.9999, threshold_cutoff=0.999, max_iterations=50, tolerance=1e-9, num_cores=1)
                    assert E.shape == (5,)
                elif what == "nmc":
                    M, E, mn = NMC(Jc, hc, mode=mode).run(20, 10, 2, 1, 1, 20, 2.0, 3, 0.01, 0.9, 0.9999, 0.999, 50, 1e-9)
                    assert mn == E.min()
                else:
                    M, E = APT_ICM(Jc, hc, mode=mode).run(np.linspace(0.3, 1.5, 3), 3, num_sweeps_MCMC=8, num_sweeps_read=4, num_swap_attempts=4,
                                                           num_swapping_pairs=1)
                    assert E.shape == (3,)
                print(f"ok    {name:24s} {mode:10s} {what}", flush=True)
            except Exception as e:
                print(f"FAIL  {name:24s} {mode:10s} {what}: {type(e).__name__}: {str(e)[:150]}", flush=True)
