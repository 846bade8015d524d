import os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "nonlocal-monte-carlo_b200"))
import numpy as np
from nlmc_b200 import NMC, _lib, nmc_core
from nlmc_b200 import instances as O  # generators of the benchmark instances
J, h = O.random_pm_graph(800, 0.06, 1)
orig = _lib.Replicas.sweep_replay
orig_phase = _lib.Replicas.set_phase
state = {}
def timed(self, perm, u, beta, *a, **k):
    t0 = time.perf_counter(); out = orig(self, perm, u, beta, *a, **k); dt = time.perf_counter() - t0
    M = out[0]
    flips = float(np.mean(M[0][1:] != M[0][:-1])) if M is not None and M.shape[1] > 1 else -1
    print(f"sweep_replay S={np.asarray(beta).shape[-1]} {dt*1e3:8.1f} ms  scaled={state.get('sc')} frozen={state.get('fr')} flip_rate={flips:.3f}", flush=True)
    return out
def phase(self, r, h_eff=None, row_scaled=None, temp_x=1.0):
    state['sc'] = None if row_scaled is None else int(np.sum(row_scaled)); state['fr'] = None if h_eff is None else int(np.sum(np.abs(h_eff) > 100))
    return orig_phase(self, r, h_eff, row_scaled, temp_x)
_lib.Replicas.sweep_replay = timed; _lib.Replicas.set_phase = phase
np.random.seed(1); random.seed(1); os.chdir("/tmp")
NMC(J, h).run(1000, 1000, 3, 1, 1, 20, 3, 3, 0.01, 0.9, 0.9999999, 0.999999, 100, np.finfo(float).eps)
