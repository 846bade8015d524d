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


# ---------------------------------------------------------------------------------------------------
# Contrived "Wishart backbone + trees" instances, restating the recipe of
# NMC/examples/contrived_wishart_backbone/contrived_instance_generator.py with the same np.random draw order, so that a
# seeded call reproduces the reference's instance files bit for bit.  Host-side data preparation, no kernel involved.
# ---------------------------------------------------------------------------------------------------
def tree_backbone_adjacency(n: int, levels: int) -> np.ndarray:
    """0/1 adjacency of a complete graph on n backbone nodes with a binary tree of `levels` levels hanging from each
    (contrived_instance_generator.py:10-46).  Children are numbered breadth first, backbone node by backbone node."""
    per_tree = 2 ** (levels + 1) - 2               # tree nodes below one backbone node
    total = n * (per_tree + 1)
    A = np.zeros((total, total))
    A[:n, :n] = 1.0 - np.eye(n)
    for i in range(n):
        base = n + i * per_tree                    # first child of backbone node i
        parents = np.concatenate(([i], base + np.arange(per_tree // 2 - 1))) if levels > 0 else np.array([], dtype=int)
        for k, p in enumerate(parents):            # parent number k (breadth first) has children base+2k, base+2k+1
            for c in (base + 2 * k, base + 2 * k + 1):
                A[p, c] = A[c, p] = 1.0
    return A


def contrived_wishart_tree(J_backbone, levels: int = 2, max_h: float = 0.2, max_outside_weight: float = 1.0,
                           max_backbone_weight: float = 10.0, num_cross_connections: int = 50,
                           max_cross_connection_weight: float = 1.0, num_remove_edges: int = 0):
    """(J, h) of one contrived instance around the planted Wishart couplings `J_backbone` (already sign-flipped, i.e. the
    `-J` of the instance file), drawing from the global np.random exactly as contrived_instance_generator.py:236-303 does:
    backbone weights (discarded later, but drawn), backbone-tree and tree-tree weights, cross connections with
    rejection, optional backbone-edge removal, then the fields."""
    Jb = J_backbone.toarray() if sp.issparse(J_backbone) else np.asarray(J_backbone, dtype=np.float64)
    b = Jb.shape[0]
    A = tree_backbone_adjacency(b, levels)
    total = len(A)
    lo_b, hi_b, lo_o, hi_o = -max_backbone_weight, max_backbone_weight, -max_outside_weight, max_outside_weight
    iu = np.triu_indices(b, 1)                      # (i, j), i < j, row-major: the order of the reference's double loop
    w = lo_b + (hi_b - lo_b) * np.random.rand(len(iu[0]))
    w = np.where((iu[0] + iu[1]) % 2 == 0, -np.abs(w), np.abs(w))
    A[iu] = w
    A[(iu[1], iu[0])] = w
    bt = (lo_o + (hi_o - lo_o) * np.random.rand(b, total - b)) * A[:b, b:]
    A[:b, b:] = bt
    A[b:, :b] = bt.T
    A[b:, b:] = (lo_o + (hi_o - lo_o) * np.random.rand(total - b, total - b)) * A[b:, b:]
    A = np.maximum(A, A.T)
    chosen = set()
    while len(chosen) < num_cross_connections:      # rejection loop: the draws depend on what was accepted before
        n1, n2 = np.random.randint(b, total), np.random.randint(b, total)
        if n1 != n2 and (n1, n2) not in chosen and (n2, n1) not in chosen:
            A[n1, n2] = A[n2, n1] = (-max_cross_connection_weight
                                     + 2 * max_cross_connection_weight * np.random.rand())
            chosen.add((n1, n2))
    removed = set()
    while len(removed) < num_remove_edges:
        n1, n2 = np.random.randint(0, b), np.random.randint(0, b)
        if n1 != n2 and A[n1, n2] != 0 and (n1, n2) not in removed and (n2, n1) not in removed:
            A[n1, n2] = A[n2, n1] = 0
            removed.add((n1, n2))
    A[:b, :b] = max_backbone_weight * Jb / np.max(np.abs(Jb))
    h = (np.random.rand(total) - 0.5) * 2 * max_h * max_backbone_weight
    return A, h


def write_instance(J, h, filename: str) -> None:
    """`i j value` text of an instance in the reference's positive-Hamiltonian convention (-J on the upper triangle
    incl. diagonal, then -h as `i i value`), contrived_instance_generator.py:211-233."""
    J = J.toarray() if sp.issparse(J) else np.asarray(J)
    with open(filename, "w") as f:
        for i, j in zip(*np.nonzero(np.triu(J))):
            f.write(f"{i} {j} {-J[i, j]}\n")
        if h is not None:
            hv = np.asarray(h).reshape(-1)
            for i in np.flatnonzero(hv):
                f.write(f"{i} {i} {-hv[i]}\n")


# ---------------------------------------------------------------------------------------------------
# Synthetic benchmark instances of the five configurations; seeded with
# np.random.RandomState(seed), couplings already normalised to max |J| = 1, h = 0.
# ---------------------------------------------------------------------------------------------------
def ea3d_pm_j(L: int, seed: int):
    """3D periodic +-J Edwards-Anderson lattice (configs C2, C4, C5): site i = x + L*(y + L*z), three forward bonds
    per site, +-1 equiprobable.  Returns (J as scipy CSR, h = zeros(N))."""
    rs = np.random.RandomState(seed)
    N = L ** 3
    idx = np.arange(N)
    x, y, z = idx % L, (idx // L) % L, idx // (L * L)
    fwd = [((x + 1) % L) + L * (y + L * z), x + L * (((y + 1) % L) + L * z), x + L * (y + L * ((z + 1) % L))]
    v = rs.choice([-1.0, 1.0], size=(3, N)).reshape(-1)
    rows, cols = np.concatenate([idx, idx, idx]), np.concatenate(fwd)
    A = sp.coo_matrix((np.concatenate([v, v]), (np.concatenate([rows, cols]), np.concatenate([cols, rows]))),
                      shape=(N, N)).tocsr()  # duplicates (L = 2) are summed, as a dense J += J.T would
    A.sum_duplicates()
    A.sort_indices()
    return A, np.zeros(N)


def random_pm_graph(N: int, p: float, seed: int):
    """Config C1: every pair i < j present with probability p, value +-1.  Returns (dense J, h = zeros(N))."""
    rs = np.random.RandomState(seed)
    iu = np.triu_indices(N, 1)
    keep = rs.rand(len(iu[0])) < p
    J = np.zeros((N, N))
    J[iu[0][keep], iu[1][keep]] = rs.choice([-1.0, 1.0], size=int(keep.sum()))
    J += J.T
    return J, np.zeros(N)


def sk_gaussian(N: int, seed: int):
    """Config C3: Sherrington-Kirkpatrick, J_ij ~ N(0, 1) / sqrt(N), symmetric, zero diagonal (not normalised)."""
    rs = np.random.RandomState(seed)
    iu = np.triu_indices(N, 1)
    J = np.zeros((N, N))
    J[iu] = rs.randn(len(iu[0])) / np.sqrt(N)
    J += J.T
    return J, np.zeros(N)
