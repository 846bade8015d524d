"""Instance I/O: the `i j value` text formats of the reference's example
scripts, parsed straight to scipy CSR (the reference goes dict -> dense N x N -> csr_matrix, which needs
N^2 memory).  Conventions follow the reference parsers line by line:

    wishart             0-based, diagonal lines skipped, h = 0          NMC/examples/wishart_example.py:8-47
    DCL                 0-based, diagonal lines skipped, h = 0          NMC/examples/DCL_example.py:8-47
    contrived wishart   0-based, diagonal lines are the field h         NMC/examples/contrived_wishart_example.py:8-57
    chimera droplet     1-based, diagonal lines are the field h         NMC/examples/chimera_example.py:8-40

Lines that are empty or start with '#' are ignored; a later line for the same pair overrides an earlier one
(dict semantics) and sets both (i,j) and (j,i).  The example scripts then flip signs to match the
Hamiltonian E = -(m^T J m/2 + m^T h): `J = -J` (all four) and `h = -h` (chimera, contrived).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def read_instance(path: str, index_base: int = 0, diagonal: str = "skip", flip_sign: bool = False):
    """Returns (J as scipy CSR [N x N], h as float64 [N, 1]).  diagonal: 'skip' or 'field'."""
    if diagonal not in ("skip", "field"):
        raise ValueError("diagonal must be 'skip' or 'field'")
    pairs = {}
    fields = {}
    with open(path, "r") as f:
        for line in f:
            line = line.strip()
            if not line or line.startswith("#"):
                continue
            x = line.split()
            i, j, v = int(float(x[0])) - index_base, int(float(x[1])) - index_base, float(x[2])
            if i == j:
                if diagonal == "field":
                    fields[i] = v
                continue
            pairs[(i, j)] = v
            pairs[(j, i)] = v
    if not pairs:
        raise ValueError(f"{path}: no couplings found")
    N = max(max(k) for k in pairs) + 1
    rows = np.fromiter((k[0] for k in pairs), dtype=np.int64, count=len(pairs))
    cols = np.fromiter((k[1] for k in pairs), dtype=np.int64, count=len(pairs))
    vals = np.fromiter(pairs.values(), dtype=np.float64, count=len(pairs))
    keep = vals != 0  # csr_matrix(dense) drops explicit zeros
    J = sp.csr_matrix((vals[keep], (rows[keep], cols[keep])), shape=(N, N))
    J.sort_indices()
    h = np.zeros((N, 1))
    for i, v in fields.items():
        if i < N:
            h[i, 0] = v
    if flip_sign:
        J, h = -J, -h
    return J, h


def read_wishart(path, flip_sign=True):
    return read_instance(path, 0, "skip", flip_sign)


def read_dcl(path, flip_sign=True):
    return read_instance(path, 0, "skip", flip_sign)


def read_contrived_wishart(path, flip_sign=True):
    return read_instance(path, 0, "field", flip_sign)


def read_chimera_droplet(path, flip_sign=True):
    return read_instance(path, 1, "field", flip_sign)
