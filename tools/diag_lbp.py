import sys, os, random
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/nonlocal-monte-carlo_b200'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from conftest import golden
from nlmc_b200 import _lib, host, nmc_core
from oracle import oracle as O
EPS=np.finfo(float).eps
for name, beta in (("npt_run_c2", 3.0), ("nmc_run_c1", 3.0), ("lbp", 3.0)):
    g = golden(name)
    J, h = g["J"], g["h"]
    norm = np.max(np.abs(J)); J = J/norm; h = h/norm
    prob = host.Problem(J, h); csr = O.Csr(J)
    lbp = _lib.Lbp(prob.inst)
    rs = np.random.RandomState(1)
    for trial in range(3):
        ms = rs.choice([-1.0,1.0], size=csr.n)
        lbp.reset(ms)
        eps_o = np.abs(h) + O._pairwise_rowsum_abs(csr)
        u = np.ascontiguousarray(csr.val*ms[csr.ci]); hm=np.zeros_like(u); tot=np.zeros(csr.n)
        lam=3.0; out=[]
        while lam >= 0.01:
            mo, io = O.lbp(csr, np.ascontiguousarray(h+lam*ms*eps_o), beta, u, hm, tot, EPS, 100)
            mg, ig = lbp.step(lam, beta, EPS, 100)
            out.append((round(lam,4), io, ig, float(np.max(np.abs(mo-mg)))))
            if io == 99 and ig == 99: break
            lam*=0.9
        print(name, trial, out)
