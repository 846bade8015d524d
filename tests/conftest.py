"""pytest configuration: the `gpu` marker and import paths.

`-m "not gpu"`: oracle vs golden vectors / live reference, host logic, C-ABI symbol checks.
`-m gpu`      : parity tests proper -- the CUDA path (through the C-ABI) vs the oracle and the goldens.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "nonlocal-monte-carlo_b200")
for p in (ROOT, PKG_DIR):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """Build libnlmc_b200.so when it is missing or older than its sources and nvcc is at hand (the snapshot sent to the
    GPU box normally carries the built library; the product itself never builds or falls back -- it fails loudly)."""
    import shutil
    import subprocess
    if shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"):
        return
    lib = os.path.join(PKG_DIR, "nlmc_b200", "libnlmc_b200.so")
    src_dir = os.path.join(PKG_DIR, "csrc")
    srcs = [os.path.join(src_dir, f) for f in os.listdir(src_dir)] + [os.path.join(ROOT, "include", "nlmc_b200.h")]
    if os.path.exists(lib) and os.path.getmtime(lib) >= max(os.path.getmtime(f) for f in srcs):
        return
    subprocess.run([sys.executable, os.path.join(PKG_DIR, "build.py")], check=False)


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


@pytest.fixture
def tmp_cwd(tmp_path, monkeypatch):
    """run() methods write PNG/NPY side-effect files into the cwd."""
    monkeypatch.chdir(tmp_path)
    return tmp_path
