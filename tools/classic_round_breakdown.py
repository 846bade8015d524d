"""Where a round of the classic (one GPU owns every slot) C5 handle spends its time: sweeps, energies, bit exchange.
python tools/classic_round_breakdown.py"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nonlocal-monte-carlo_b200"))
from nlmc_b200 import _lib, host, instances  # noqa: E402

A, h = instances.ea3d_pm_j(64, 5)
prob = host.Problem(A, h)
betas = np.linspace(0.2, 2.0, 32)
stream = torch.cuda.Stream()


def timed(fn, reps=10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        for _ in range(3):
            fn()
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


msc = _lib.Msc(prob.inst, betas, 128, seed=1)
msc.set_stream(stream.cuda_stream)
E = torch.zeros((32, 128), dtype=torch.float64, device="cuda")
out = {"sweeps16_ms": timed(lambda: msc.sweep(16)), "energies_ms": timed(lambda: msc.energies_into(E)),
       "round16_ms": timed(lambda: msc.round(16, 10)), "round0_ms": timed(lambda: msc.round(0, 10))}
print(json.dumps(out), flush=True)
msc.close()
