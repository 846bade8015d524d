"""Time-to-target harness (BASELINE.json metric "NPT time-to-target-energy vs CPU ref").

Instance: the reference's Chimera droplet instance 001 (128 spins, real J, fields) whose ground-state energy
ships with the reference (golden copy in tests/golden/known_answer_chimera128.npz).  Target: the ground state.
Both arms run the same algorithm -- parallel tempering over the same beta ladder, `spm` heat-bath sweeps per
round, adjacent swaps with min(1, exp(dB*dE)) -- until the best energy reaches the target:
  GPU  : dense tensor-core engine (K3), `runs` independent ladders at once, swaps as beta-label exchanges;
  CPU  : the oracle C port of MCMC (reference algorithm) driven from Python, one ladder.
Prints one JSON line per arm with the median over `repeats` seeds.
"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "nonlocal-monte-carlo_b200"))
from nlmc_b200 import _lib, host, instances  # noqa: E402
from oracle import oracle as O  # noqa: E402  (CPU arm)

g = np.load(os.path.join(ROOT, "tests", "golden", "known_answer_chimera128.npz"))
with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
    f.write(str(g["instance_text"]))
J, h = instances.read_chimera_droplet(f.name)
norm = abs(J).max()
Jn, hn = J / norm, h.reshape(-1) / norm
target = float(g["gs_energy"]) / norm
n = Jn.shape[0]
betas = np.geomspace(0.5, 12.0, 16)
spm, max_rounds, tol = 5, 4000, 1e-6


def gpu_arm(seed, runs=64):
    prob = host.Problem(Jn, hn)
    nb = len(betas)
    lab = np.tile(np.arange(nb), runs)          # beta label of every row (row = run*nb + slot)
    d = _lib.Dense(prob.inst, betas[lab], n_split=3, seed=seed)
    rs = np.random.RandomState(seed)
    d.sweep(1); d.energies()                      # warm-up (graph capture)
    d.set_spins(rs.choice([-1, 1], size=(runs * nb, n)).astype(np.int8))
    t0 = time.perf_counter()
    for rnd in range(1, max_rounds + 1):
        d.sweep(spm)
        E = d.energies()
        if E.min() <= target + 1e-4:              # fp32 field GEMM: confirm exactly below
            S = d.get_spins()
            if prob.inst.energy_states(S[[int(np.argmin(E))]])[0] <= target + tol:
                return time.perf_counter() - t0, rnd * spm
        # adjacent exchanges, even/odd alternation, as label swaps
        order = np.argsort(lab.reshape(runs, nb), axis=1)          # order[run][b] = slot holding beta b
        Eb = np.take_along_axis(E.reshape(runs, nb), order, axis=1)
        for b in range(rnd % 2, nb - 1, 2):
            acc = rs.rand(runs) < np.minimum(1.0, np.exp((betas[b + 1] - betas[b]) * (Eb[:, b + 1] - Eb[:, b])))
            lo, hi = order[:, b].copy(), order[:, b + 1].copy()
            rows = np.arange(runs)[acc]
            lab2 = lab.reshape(runs, nb)
            lab2[rows, lo[acc]], lab2[rows, hi[acc]] = b + 1, b
        d.set_betas(betas[lab])
    return float("inf"), max_rounds * spm


def cpu_arm(seed):
    csr = O.Csr(Jn)
    rs = np.random.RandomState(seed)
    nb = len(betas)
    S = rs.choice([-1, 1], size=(nb, n)).astype(np.int8)
    t0 = time.perf_counter()
    for rnd in range(1, max_rounds + 1):
        for b in range(nb):
            _, S[b] = O.mcmc(csr, hn, S[b], np.full(spm, betas[b]), rng=rs, use_lut=False)
        E = O.energy(csr, hn, S)
        if E.min() <= target + tol:
            return time.perf_counter() - t0, rnd * spm
        for b in range(rnd % 2, nb - 1, 2):
            if rs.rand() < min(1.0, np.exp((betas[b + 1] - betas[b]) * (E[b + 1] - E[b]))):
                S[[b, b + 1]] = S[[b + 1, b]]
                E[[b, b + 1]] = E[[b + 1, b]]
    return float("inf"), max_rounds * spm


if __name__ == "__main__":
    repeats = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    for name, arm in (("gpu dense engine, 64 ladders x 16 betas", gpu_arm), ("cpu oracle port, 1 ladder x 16 betas, 1 core", cpu_arm)):
        res = [arm(100 + i) for i in range(repeats)]
        ts = sorted(r[0] for r in res)
        print(json.dumps({"what": "time to ground state, Chimera droplet 128 inst 001", "arm": name,
                          "target_energy_normalised": target, "median_seconds": ts[len(ts) // 2],
                          "all_seconds": ts, "sweeps_per_ladder_median": sorted(r[1] for r in res)[len(res) // 2]}), flush=True)
