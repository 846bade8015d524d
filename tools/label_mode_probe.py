"""Timing probe: C5-shaped rounds in the classic (bit-exchange) and the beta-label form, and blocks of a sharded ladder
(4 / 8 / 16 slots = 16 / 32 / 64 words per site row).  python tools/label_mode_probe.py [L]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nonlocal-monte-carlo_b200"))
from nlmc_b200 import _lib, host, instances  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 64
A, h = instances.ea3d_pm_j(L, 5)
prob = host.Problem(A, h)
betas = np.linspace(0.2, 2.0, 32)
n = prob.n


def time_rounds(msc, rounds=10, spm=16, pairs=10, sweeps_only=False):
    for _ in range(3):
        msc.sweep(spm) if sweeps_only else msc.round(spm, pairs)
    msc.sync()
    msc.timer_mark(0)
    for _ in range(rounds):
        msc.sweep(spm) if sweeps_only else msc.round(spm, pairs)
    msc.timer_mark(1)
    msc.sync()
    ms = msc.timer_elapsed_ms() / rounds
    return ms, msc.n_beta * msc.n_ladders * n * spm / (ms * 1e-3)


for name, kw in [("classic 32 slots", dict()), ("labelled 32 slots", dict(labelled=True)),
                 ("labelled block of 16", dict(labelled=True, slot_begin=8, slot_count=16)),
                 ("labelled block of 8", dict(labelled=True, slot_begin=8, slot_count=8)),
                 ("labelled block of 4", dict(labelled=True, slot_begin=8, slot_count=4))]:
    msc = _lib.Msc(prob.inst, betas, 128, seed=1, **kw)
    whole = msc.n_beta == 32
    ms, rate = time_rounds(msc, sweeps_only=not whole)
    out = {"variant": name, "L": L, "words_per_row": msc.n_words, "ms_per_16_sweeps": ms, "attempts_per_s": rate}
    if whole:
        ms2, rate2 = time_rounds(msc, sweeps_only=True)
        out.update(ms_sweeps_only=ms2, attempts_per_s_sweeps_only=rate2)
    print(json.dumps(out), flush=True)
    msc.close()
