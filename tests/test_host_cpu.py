"""CPU-side tests of the product: the C-ABI library loads and exports every symbol include/nlmc_b200.h
declares, the host logic matches the reference's, and the product fails loudly without a GPU."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "nlmc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nlmc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from nlmc_b200 import _lib
    L = _lib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/nlmc_b200.h but not exported"
    assert sorted(_lib.exported_symbols()) == declared, "ctypes binding table out of sync with the header"
    assert L.nlmc_version() >= 100


def test_no_cpu_fallback():
    """Without a CUDA device the product must raise, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from nlmc_b200 import NPT, _lib
    J = np.array([[0.0, 1.0], [1.0, 0.0]])
    with pytest.raises(_lib.NlmcError):
        NPT(J, np.zeros(2)).run(np.array([0.5, 1.0]), 2, [False, False], num_sweeps_MCMC=4, num_sweeps_read=4,
                                num_swap_attempts=2)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "nonlocal-monte-carlo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "nlmc_oracle" not in text, f


def test_beta_schedule_matches_reference_rule():
    from nlmc_b200 import host
    from oracle import oracle as O
    for n, b in ((0, 1.0), (1, 2.0), (7, 3.0), (100, 2.5)):
        assert np.array_equal(host.beta_schedule(n, b, True, 1, 0), O.anneal_schedule(n, b, True, 1, 0))
        assert np.array_equal(host.beta_schedule(n, b), np.full(n, b))
    with pytest.raises(ValueError):
        host.beta_schedule(-100, 1.0)


def test_pair_selection_and_errors():
    import random
    from nlmc_b200 import host
    from oracle import oracle as O
    pairs = [(i, i + 1) for i in range(1, 9)]
    random.seed(3)
    a = host.select_non_overlapping_pairs(pairs, 3)
    random.seed(3)
    b = O.select_non_overlapping_pairs(pairs, 3)
    assert a == b and len({x for p in a for x in p}) == 6
    with pytest.raises(ValueError, match="Cannot find non-overlapping pairs"):
        host.select_non_overlapping_pairs([(1, 2)], 2)


def test_constructor_contract():
    """Constructor attribute contract of the four classes."""
    from nlmc_b200 import APT_ICM, NMC, NPT, APT_preprocessor
    J = np.array([[0.0, 2.0], [2.0, 0.0]])
    h = np.array([[0.5], [-0.5]])
    for cls in (NMC, NPT):
        o = cls(J, h)
        assert o.J is J and o.h.shape == (2,)
    for cls in (APT_preprocessor, APT_ICM):
        o = cls(J, [0.5, -0.5])
        assert o.h.shape == (2, 1)
    assert APT_preprocessor(J, h).N == 2


def test_find_clusters_matches_oracle():
    from nlmc_b200 import nmc_core
    from oracle import oracle as O
    J, h = O.random_pm_graph(40, 0.2, 9)
    csr = O.Csr(J)

    class P:  # the host part of host.Problem, without the device
        n, rp, ci, val = csr.n, csr.rp, csr.ci, csr.val

        def neighbours(self, i):
            b, e = self.rp[i], self.rp[i + 1]
            return self.ci[b:e][self.val[b:e] != 0]

    rs = np.random.RandomState(0)
    for thr_i, thr_c in ((0.9, 0.89), (0.9, 0.5), (0.99, 0.8)):
        marg = np.tanh(rs.randn(40) * 2)
        a = nmc_core.find_clusters(P(), marg, thr_i, thr_c, 0.01)
        b = O.find_clusters(csr, marg, thr_i, thr_c, 0.01)
        assert len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b))


def test_bench_module_is_whole_and_reference_arm_runs():
    """bench.py keeps every piece of its contract (a past edit once dropped the clock sampler) and its reference arm
    -- the oracle port on the host cores -- prints the contract's JSON line on a tiny workload."""
    import importlib.util
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(root, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for name in ("ClockSampler", "ea3d_csr", "measured_peaks", "ncu_traffic", "cpu_port_rate", "run_reference",
                 "workload_config", "run_ours", "main"):
        assert hasattr(mod, name), name
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--L", "8", "--n-beta", "4",
                          "--steps", "1", "--warmup", "1", "--ref-sweeps", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 0


def test_result_cache_never_hands_out_a_buffer_that_is_still_referenced():
    """The caching allocator behind the large float64 results (NPT.run's M): a buffer is reused only when neither the array
    handed out earlier nor a view of it is alive."""
    from nlmc_b200 import _lib
    _lib.release_host_cache()
    shape = (1 << 20, 16)            # 128 MB: above the caching threshold
    a = _lib.result_cache.take(shape)
    a[0, 0] = 42.0
    addr_a = a.ctypes.data
    b = _lib.result_cache.take(shape)      # a is alive: a different buffer
    assert b.ctypes.data != addr_a and a[0, 0] == 42.0
    addr_b = b.ctypes.data
    v = b[:10]                             # a view keeps the buffer busy
    del b
    c = _lib.result_cache.take(shape)
    assert c.ctypes.data != addr_b
    addr_c = c.ctypes.data
    del c, v
    d = _lib.result_cache.take((1 << 19, 16))   # free and large enough: reused
    assert d.ctypes.data == addr_c and d.shape == (1 << 19, 16) and d.dtype == np.float64
    small = _lib.result_cache.take((4, 4))      # small results are plain arrays
    assert small.shape == (4, 4)
    del d
    _lib.release_host_cache()
