import os, sys, json
import numpy as np, torch
ROOT = "/root/repo"
sys.path.insert(0, os.path.join(ROOT, "nonlocal-monte-carlo_b200"))
from nlmc_b200 import _lib, host, instances
A, h = instances.ea3d_pm_j(64, 5)
prob = host.Problem(A, h)
betas = np.linspace(0.2, 2.0, 32)
stream = torch.cuda.Stream()
def timed(fn, reps=10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        for _ in range(3): fn()
        e0.record(stream)
        for _ in range(reps): fn()
        e1.record(stream)
    e1.synchronize()
    return e0.elapsed_time(e1) / reps
for kw in ({}, dict(labelled=True, slot_begin=0, slot_count=4)):
    msc = _lib.Msc(prob.inst, betas, 128, seed=1, **kw)
    msc.set_stream(stream.cuda_stream)
    msc.sweep(4)
    E = torch.zeros((msc.n_beta, 128), dtype=torch.float64, device="cuda")
    ref = None
    for ch, th, ct in (("", "", ""), ("", "128", "592"), ("", "128", "296"), ("", "256", "592"), ("", "256", "296"), ("", "256", "148"), ("32", "256", "296"), ("", "512", "148"), ("32", "512", "148")):
        for k, v in (("NLMC_ENERGY_CHUNK", ch), ("NLMC_ENERGY_THREADS", th), ("NLMC_ENERGY_CTAS", ct)):
            if v: os.environ[k] = v
            else: os.environ.pop(k, None)
        ch = f"{ch}/{th}/{ct}"
        t = timed(lambda: msc.energies_into(E))
        torch.cuda.synchronize()
        s = float(E.sum().item())
        if ref is None: ref = s
        print(kw.get("slot_count", 32), "chunk", ch or "auto", f"{t*1e3:.1f} us", "sum ok" if s == ref else "SUM DIFFERS")
    os.environ.pop("NLMC_ENERGY_CHUNK", None)
    msc.close()
