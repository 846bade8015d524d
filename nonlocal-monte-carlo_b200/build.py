"""Builds libnlmc_b200.so (hand-written sm_100a CUDA + the C ABI of include/nlmc_b200.h) in-tree.

    python nonlocal-monte-carlo_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "nlmc_b200", "libnlmc_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", CSRC]
FLAGS += os.environ.get("NLMC_NVCC_EXTRA", "").split()  # e.g. -DNLMC_PHILOX_ROUNDS=7 for an experiment build


CXX = os.environ.get("CXX", "g++")
CXXFLAGS = ["-O3", "-std=c++17", "-fPIC", "-pthread", "-I", os.path.join(ROOT, "include"), "-I", CSRC]


def sources():
    """CUDA sources (nvcc, sm_100a) and the plain C++ host helpers (g++)."""
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "nlmc_b200.h"))
    jobs = []
    objs = []
    for src in sources():
        obj = os.path.join(OBJ, os.path.splitext(os.path.basename(src))[0] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            if src.endswith(".cpp"):
                cmd = [CXX, *CXXFLAGS, "-c", src, "-o", obj]
            else:
                cmd = [NVCC, *FLAGS, "-c", src, "-o", obj]
                if verbose:
                    cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)

    def run(cmd):
        p = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, p.returncode, p.stdout + p.stderr

    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for cmd, rc, out in ex.map(run, jobs):
                if verbose or rc:
                    print(" ".join(cmd))
                    print(out)
                if rc:
                    raise RuntimeError(f"nvcc failed for {cmd[-3]}")
    if jobs or force or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB, *objs, "-lcudart", "-lpthread"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode:
            print(p.stdout + p.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
