"""Stage timing of config C5 through the pieces NPT(J, h, mode='production').run is made of (one GPU).
    python tools/c5_api_stages.py [rounds]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nonlocal-monte-carlo_b200"))
from nlmc_b200 import _lib, host, instances  # noqa: E402

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 20
A, h = instances.ea3d_pm_j(64, 5)
betas = np.linspace(0.2, 2.0, 32)
n = A.shape[0]


def once(tag):
    T = [time.perf_counter()]
    def lap(name):
        T.append(time.perf_counter())
        print(f"  {tag} {name:28s} {1e3 * (T[-1] - T[-2]):8.2f} ms", flush=True)
    prob = host.Problem(A, h)
    lap("host.Problem (CSR upload)")
    msc = _lib.Msc(prob.inst, betas, 128, 1234)
    lap("Msc create")
    msc.sync()
    lap("sync after create")
    for _ in range(rounds - 1):
        msc.round(16, 10)
    lap("enqueue rounds")
    msc.sync()
    lap("wait rounds")
    M = _lib.result_cache.take((32 * n, 16))
    lap("M from the result cache")
    msc.sweep_record_f64(16, ladder=0, out=M)
    lap("record round + copy + widen")
    msc.round(0, 10)
    c = msc.swap_counts(rounds)
    lap("last exchange + counts")
    msc.close()
    lap("Msc close")
    prob.inst.close()
    lap("Instance close")
    print(f"  {tag} total {1e3 * (T[-1] - T[0]):.2f} ms")
    return M


for t in range(3):
    M = once(f"call{t}")
    del M
