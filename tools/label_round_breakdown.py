"""Where a round of a beta-sharded ladder spends its time on ONE rank (no NCCL): sweeps, bit-sliced energies, label exchange
+ threshold planes, for blocks of 32 / 16 / 8 / 4 slots of the C5 ladder.  python tools/label_round_breakdown.py"""
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


def timed(fn, reps=20):
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


for count in (32, 16, 8, 4):
    msc = _lib.Msc(prob.inst, betas, 128, seed=1, labelled=True, slot_begin=0, slot_count=count)
    msc.set_stream(stream.cuda_stream)
    E_full = torch.zeros((32, 128), dtype=torch.float64, device="cuda")
    E_full += torch.linspace(-4e5, -1e5, 32, device="cuda", dtype=torch.float64)[:, None]
    out = {"slots": count, "words_per_row": msc.n_words,
           "sweeps16_ms": timed(lambda: msc.sweep(16)),
           "energies_ms": timed(lambda: msc.energies_into(E_full[:count])),
           "exchange_ms": timed(lambda: msc.exchange_labels_from(E_full, 10)),
           "allgather_like_copy_ms": timed(lambda: E_full.copy_(E_full.clone()))}
    print(json.dumps(out), flush=True)
    msc.close()
