"""Dev tool: time K1 (exact replay) on the C1 instance in its regimes; optional ncu target."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "nonlocal-monte-carlo_b200"))
import numpy as np
from nlmc_b200 import _lib, host
from nlmc_b200 import instances as O  # generators of the benchmark instances
J, h = O.random_pm_graph(800, 0.06, 1)
prob = host.Problem(J, h)
n, S = 800, 200
rs = np.random.RandomState(0)
perm = np.stack([rs.permutation(n) for _ in range(S)]).astype(np.int32)[None]
u = rs.rand(1, S, n)
m0 = rs.choice([-1, 1], size=(1, n)).astype(np.int8)
reps = _lib.Replicas(prob.inst, 1, m0)
in_cl = rs.rand(n) < 0.15
def run(tag, beta, h_eff=None, scaled=None, lut=True):
    reps.set_phase(0, h_eff, scaled, 20.0)
    reps.set_spins(m0)
    sched = np.full((1, S), beta)
    L = prob.tanh_lut(sched) if lut else None
    reps.sweep_replay(perm, u, sched, L, prob.lut_half if lut else 0)
    t0 = time.perf_counter()
    reps.sweep_replay(perm, u, sched, L, prob.lut_half if lut else 0)
    dt = time.perf_counter() - t0
    print(f"{tag:28s} {dt*1e3:8.1f} ms  {dt/(S*n)*1e9:7.1f} ns/attempt")
run("plain beta=3 (LUT)", 3.0)
run("plain beta=0.3 (LUT)", 0.3)
run("plain beta=3 (no LUT)", 3.0, lut=False)
he = np.zeros(n); he[~in_cl] = m0[0][~in_cl] * 10000.0
run("phase C (scaled+frozen)", 3.0, he, in_cl.astype(np.uint8))
he = np.zeros(n); he[in_cl] = m0[0][in_cl] * 10000.0
run("phase NC (frozen backbone)", 3.0, he, None)
os.environ["NLMC_REPLAY_GENERAL"] = "1"
run("general kernel, plain beta=3", 3.0)
