"""Small fixed workload for ncu captures of the colour-sweep kernel at C5 size (L = 64, 32 betas x 128 ladders):
    python tools/profile_msc.py classic|labelled|block4 [sweeps]
classic = scalar thresholds (one GPU owns every slot), labelled = bit planes (beta-label exchange), block4 = a block of 4
slots of a sharded ladder (16 words per site row, several sites per warp)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nonlocal-monte-carlo_b200"))
from nlmc_b200 import _lib, host, instances  # noqa: E402

variant = sys.argv[1] if len(sys.argv) > 1 else "classic"
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
A, h = instances.ea3d_pm_j(64, 5)
prob = host.Problem(A, h)
betas = np.linspace(0.2, 2.0, 32)
kw = {"classic": {}, "labelled": dict(labelled=True), "block4": dict(labelled=True, slot_begin=8, slot_count=4)}[variant]
msc = _lib.Msc(prob.inst, betas, 128, seed=1, **kw)
os.environ["NLMC_MSC_GRAPHS"] = "0"
msc.sweep(sweeps)
msc.sync()
print(variant, "done", msc.n_words, "words per row")
msc.close()
