"""Probe of the end-to-end pipeline of bench.py: how many alternating handles hide the host copies."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "nonlocal-monte-carlo_b200"))
import numpy as np, torch
from nlmc_b200 import _lib, host
from bench import ea3d_csr
A = ea3d_csr(64, 5)
prob = host.Problem(A, np.zeros(64**3))
betas = np.linspace(0.2, 2.0, 32)
spm, pairs, steps = 16, 10, 24
for nh in (1, 2, 3, 4):
    for copies in (True, False):
        hs = [_lib.Msc(prob.inst, betas, 128, seed=100 + i) for i in range(nh)]
        shape = hs[0].packed_shape()
        hin = [torch.empty(shape, dtype=torch.int32, pin_memory=True) for _ in range(nh)]
        hout = [torch.empty(shape, dtype=torch.int32, pin_memory=True) for _ in range(nh)]
        hE = [torch.empty((32, 128), dtype=torch.float64, pin_memory=True) for _ in range(nh)]
        for i in range(nh):
            hs[i].get_packed(hin[i].numpy().view(np.uint32)); hout[i].copy_(hin[i])
        def run(count):
            for i in range(count):
                k = i % nh
                hs[k].sync()
                hin[k], hout[k] = hout[k], hin[k]
                if copies:
                    hs[k].round_host_async(hin[k].data_ptr(), spm, pairs, hout[k].data_ptr(), hE[k].data_ptr())
                else:
                    hs[k].round_host_async(None, spm, pairs, None, None)
            for h in hs: h.sync()
        run(2 * nh)
        t0 = time.perf_counter(); run(steps); dt = time.perf_counter() - t0
        print(f"handles={nh} copies={copies}: {1e3 * dt / steps:.3f} ms/step", flush=True)
        for h in hs: h.close()
