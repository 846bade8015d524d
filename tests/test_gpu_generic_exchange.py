"""GPU: replica exchange of the generic production engines (K2a sparse, K3 dense) as a device-side permutation of beta
labels (nlmc_col_exchange / nlmc_dense_exchange, csrc/nlmc_exchange.cuh; NPT/npt.py:649-680 in the label form of
SURVEY D4), and NPT.run on top of it: nothing but the last round's record leaves the GPU.

  * labels stay permutations per ladder, exchanges happen, per-round counts are reported;
  * with exchanges the per-temperature mean energy is still the exact Boltzmann one (full enumeration), 3 sigma;
  * NPT.run(production) returns Energy[r] = min over the first spr columns of the energies of the returned M."""
import itertools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nl():
    from nlmc_b200 import _lib, host
    return type("NL", (), dict(lib=_lib, host=host))


def _small_instance(n, seed):
    rs = np.random.RandomState(seed)
    iu = np.triu_indices(n, 1)
    keep = rs.rand(len(iu[0])) < 0.45
    J = np.zeros((n, n))
    J[iu[0][keep], iu[1][keep]] = rs.randn(int(keep.sum()))
    J += J.T
    J /= np.max(np.abs(J))
    return J, rs.randn(n) * 0.2


@pytest.mark.parametrize("engine", ["Col", "Dense"])
def test_label_exchange_samples_exact_boltzmann(nl, engine):
    from oracle import oracle as O
    n = 10
    J, h = _small_instance(n, 21)
    states = np.array(list(itertools.product([-1, 1], repeat=n)), dtype=np.int8)
    E_all = O.energy(O.Csr(J), h, states)
    betas = np.array([0.3, 0.7, 1.2, 1.9])
    ladders = 512
    prob = nl.host.Problem(J, h)
    rows = np.tile(betas, ladders)
    d = nl.lib.Col(prob.inst, rows, seed=5) if engine == "Col" else nl.lib.Dense(prob.inst, rows, n_split=3, seed=5)
    d.ladders(betas)
    rs = np.random.RandomState(1)
    d.set_spins(rs.choice([-1, 1], size=(len(rows), n)).astype(np.int8))
    for _ in range(60):
        d.sweep(3)
        d.exchange(2)
    T = 150
    acc = np.zeros((len(betas), ladders))
    for _ in range(T):
        d.sweep(3)
        d.exchange(2)
        lab, _ = d.labels(0)
        E = prob.inst.energy_states(d.get_spins())              # fp64 energies of the states (K4)
        by_beta = np.empty((ladders, len(betas)))
        np.put_along_axis(by_beta, lab.reshape(ladders, -1).astype(np.int64), E.reshape(ladders, -1), axis=1)
        acc += by_beta.T
    lab, counts = d.labels(T)
    assert np.array_equal(np.sort(lab.reshape(ladders, -1), axis=1), np.tile(np.arange(4), (ladders, 1)))
    assert counts.shape == (T,) and counts.min() > 0
    assert not np.array_equal(lab.reshape(ladders, -1), np.tile(np.arange(4), (ladders, 1)))
    per_ladder = acc / T
    for b, beta in enumerate(betas):
        w = np.exp(-beta * (E_all - E_all.min()))
        w /= w.sum()
        exact = (w * E_all).sum()
        mean, err = per_ladder[b].mean(), per_ladder[b].std(ddof=1) / np.sqrt(ladders)
        assert abs(mean - exact) <= 3.0 * err + 1e-9, (engine, beta, mean, exact, err)
    d.close()


@pytest.mark.parametrize("kind", ["sparse_real", "dense_gauss", "pm_graph"])
def test_npt_run_production_generic_engines(nl, kind, tmp_cwd):
    """NPT.run(mode='production') without NMC replicas on a non-lattice instance: device-side label exchange; the
    returned Energy is the minimum over the first spr columns of the fp64 energies of the returned M (NPT/npt.py:686-692),
    rows in beta order (colder rows have lower mean energy), several independent runs side by side."""
    from nlmc_b200 import NPT
    from oracle import oracle as O
    if kind == "sparse_real":
        J, h = _small_instance(48, 4)
    elif kind == "dense_gauss":
        J, h = O.sk_gaussian(160, 9)
    else:
        J, h = O.random_pm_graph(120, 0.08, 3)
    J = np.asarray(J.todense()) if hasattr(J, "todense") else np.asarray(J)
    n = J.shape[0]
    R = 6
    betas = np.linspace(0.2, 2.5, R)
    np.random.seed(3)
    obj = NPT(J, h, mode="production")
    obj.num_runs = 3
    M, E = obj.run(betas, R, [False] * R, num_sweeps_MCMC=400, num_sweeps_read=200, num_swap_attempts=20,
                   num_swapping_pairs=2)
    spm, spr = 20, 10
    assert M.shape == (R * n, spm) and E.shape == (R,) and set(np.unique(M)) <= {-1.0, 1.0}
    assert obj.energies_all_runs.shape == (R, 3)
    norm = np.max(np.abs(J))
    csr = O.Csr(J / norm)
    means = []
    for r in range(R):
        Er = O.energy(csr, np.asarray(h).reshape(-1) / norm, M[r * n:(r + 1) * n].T.astype(np.int8))
        np.testing.assert_allclose(E[r], Er[:spr].min(), rtol=1e-9, atol=1e-9)
        means.append(Er.mean())
    assert means[-1] < means[0]                       # rows are in beta order: the coldest is the lowest
    assert np.all(np.diff(means) < 0.15 * abs(means[-1]))   # and roughly monotone in between


def test_hybrid_ladder_matches_single_engine_ladder(nl, tmp_cwd, monkeypatch):
    """Config C2's shape on a small lattice: NPT.run with NMC on the 2 coldest of 6 replicas.  The hybrid path (plain replicas
    bit-packed on K2, NMC replicas on K2a, exchanges between the two kinds) and the single-engine path (everything on K2a)
    sample the same ladder: per-replica mean of the returned last-round energies over independent runs within 3 sigma
    (two-sample), and the hybrid's returned energies belong to its returned states."""
    from nlmc_b200 import NPT
    from oracle import oracle as O
    A, h = O.ea3d_pm_j(4, 6)
    n, R = 64, 6
    csr = O.Csr(A)
    betas = np.linspace(0.3, 1.6, R)
    kw = dict(num_sweeps_MCMC=240, num_sweeps_read=240, num_swap_attempts=8, num_swapping_pairs=2, num_cycles=2,
              global_beta=2.5, lambda_start=3.0, lambda_end=0.01, threshold_initial=0.9999, threshold_cutoff=0.999,
              max_iterations=100, tolerance=1e-9)
    doNMC = [False] * 4 + [True] * 2

    def sample(runs, seed0):
        out = []
        for k in range(runs):
            np.random.seed(seed0 + k)
            import random
            random.seed(seed0 + k)
            M, E = NPT(A, h, mode="production").run(betas, R, doNMC, **kw)
            spm = M.shape[1]
            Er = np.stack([O.energy(csr, h, M[r * n:(r + 1) * n].T.astype(np.int8)) for r in range(R)])
            assert np.array_equal(E, Er.min(axis=1)) and spm == 30
            out.append(Er.mean(axis=1))
        return np.array(out)

    hyb = sample(30, 100)
    monkeypatch.setenv("NLMC_NO_HYBRID", "1")
    ref = sample(30, 500)
    for r in range(R):
        err = np.sqrt(hyb[:, r].var(ddof=1) / len(hyb) + ref[:, r].var(ddof=1) / len(ref))
        assert abs(hyb[:, r].mean() - ref[:, r].mean()) <= 3.0 * err + 1e-9, (r, hyb[:, r].mean(), ref[:, r].mean(), err)
