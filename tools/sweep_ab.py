"""A/B timing of the colour-sweep kernel at C5 size (CUDA events on the handle's stream):
    [NLMC_LIB_PATH=variant.so] python tools/sweep_ab.py [classic|labelled|block4 ...]
Prints ms per sweep, attempts/s and a checksum of the packed state (equal checksums = identical trajectories)."""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nonlocal-monte-carlo_b200"))
from nlmc_b200 import _lib, host, instances  # noqa: E402

variants = sys.argv[1:] or ["classic"]
A, h = instances.ea3d_pm_j(64, 5)
prob = host.Problem(A, h)
betas = np.linspace(0.2, 2.0, 32)
KW = {"classic": {}, "labelled": dict(labelled=True), "block4": dict(labelled=True, slot_begin=8, slot_count=4),
      "block16": dict(labelled=True, slot_begin=0, slot_count=16)}
for v in variants:
    msc = _lib.Msc(prob.inst, betas, 128, seed=1, **KW[v])
    msc.sweep(8)
    msc.sync()
    best = 1e9
    for rep in range(3):
        msc.timer_mark(0)
        msc.sweep(32)
        msc.timer_mark(1)
        best = min(best, msc.timer_elapsed_ms() / 32)
    crc = zlib.crc32(msc.get_packed().tobytes())
    att = msc.n_beta * msc.n_ladders * msc.n
    print(f"{os.path.basename(_lib.LIB_PATH)} {v}: {best:.4f} ms/sweep  {att / best * 1e3:.4g} attempts/s  crc {crc:08x}", flush=True)
    msc.close()
