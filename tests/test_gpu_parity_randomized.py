"""Randomised parity sweep on the B200: many small, ragged instances -- n from 1 to ~90, empty rows, isolated spins,
integer and real couplings (also |J| > 1 and explicit zeros), fields or none, zeros in the start state, scaled /
frozen phases, annealed schedules -- each compared with the CPU oracle: K1 trajectories and K7 cluster labels bit for
bit, K4 energies exactly on integer instances and to 1e-9 otherwise (north_star tolerance)."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


def random_instance(rs, case):
    n = int(rs.choice([1, 2, 3, 5, 17, 31, 32, 33, 64, 90]))
    density = float(rs.choice([0.0, 0.05, 0.3, 1.0]))
    kind = ("pm1", "int", "real")[case % 3]
    U = np.triu(rs.rand(n, n) < density, 1)
    if kind == "pm1":
        V = rs.choice([-1.0, 1.0], size=(n, n))
    elif kind == "int":
        V = rs.randint(-5, 6, size=(n, n)).astype(float)        # contains zeros: dropped by csr_matrix like the reference
    else:
        V = rs.randn(n, n)
    J = np.where(U, V, 0.0)
    J = J + J.T
    h = np.zeros(n) if case % 2 == 0 else (rs.randint(-2, 3, size=n).astype(float) if kind != "real" else 0.4 * rs.randn(n))
    return J, h, kind


@pytest.mark.parametrize("block", range(6))
def test_k1_k4_random_instances(block):
    from nlmc_b200 import _lib, host
    from oracle import oracle as O
    rs = np.random.RandomState(1000 + block)
    for case in range(block * 8, block * 8 + 8):
        J, h, kind = random_instance(rs, case)
        n = J.shape[0]
        csr = O.Csr(J)
        prob = host.Problem(J, h)
        R, S = 3, int(rs.randint(1, 5))
        reps = _lib.Replicas(prob.inst, R)
        m0 = rs.choice([-1, 1], size=(R, n)).astype(np.int8)
        if n > 2:
            m0[1, rs.randint(n)] = 0
        sched = np.stack([host.beta_schedule(S, float(rs.uniform(0.2, 3.0)), anneal=bool(r == 2), sweeps_per_beta=1,
                                             initial_beta=0.1) for r in range(R)])
        perm = np.stack([np.stack([rs.permutation(n) for _ in range(S)]) for _ in range(R)]).astype(np.int32)
        u = rs.rand(R, S, n)
        h_eff, scaled = [None] * R, [None] * R
        if case % 4 == 1 and n > 1:  # an NMC phase on replica 0: some rows at beta/temp_x, the rest frozen
            in_cl = rs.rand(n) < 0.5
            he = np.asarray(h, dtype=float).copy()
            he[in_cl] /= 7.0
            he[~in_cl] = m0[0][~in_cl] * 10000.0
            h_eff[0], scaled[0] = he, in_cl
            reps.set_phase(0, he, in_cl.astype(np.uint8), 7.0)
        reps.set_spins(m0)
        M, E = reps.sweep_replay(perm, u, sched, prob.tanh_lut(sched), prob.lut_half)
        rows = csr.row_of
        for r in range(R):
            c = csr if scaled[r] is None else csr.with_values(np.where(scaled[r][rows], csr.val / 7.0, csr.val))
            Mo, _ = O.mcmc(c, h if h_eff[r] is None else h_eff[r], m0[r], sched[r], perm=perm[r], u=u[r])
            assert np.array_equal(M[r], Mo), f"case {case} ({kind}, n={n}) replica {r}"
            Eo = O.energy(csr, h, Mo)
            if kind != "real":
                assert np.array_equal(E[r], Eo), f"case {case} energies"
            else:
                np.testing.assert_allclose(E[r], Eo, rtol=1e-9, atol=1e-12)
        # K4 on arbitrary states (zeros included)
        states = rs.choice([-1, 0, 1], size=(4, n)).astype(np.int8)
        np.testing.assert_allclose(prob.inst.energy_states(states), O.energy(csr, h, states), rtol=1e-9, atol=1e-12)
        reps.close()


@pytest.mark.parametrize("block", range(3))
def test_k7_random_pairs(block):
    from nlmc_b200 import _lib, host
    from oracle import oracle as O
    rs = np.random.RandomState(2000 + block)
    for case in range(8):
        J, h, _ = random_instance(rs, case)
        n = J.shape[0]
        prob = host.Problem(J, h)
        P = 5
        s1 = rs.choice([-1, 1], size=(P, n)).astype(np.int8)
        s2 = np.where(rs.rand(P, n) < rs.choice([0.0, 0.2, 0.5, 1.0]), -s1, s1).astype(np.int8)
        labels, counts = _lib.icm_clusters(prob.inst, s1, s2)
        csr = O.Csr(J)
        for p in range(P):
            lab_o, cnt_o = O.disagreement_clusters(csr, s1[p], s2[p])
            assert counts[p] == cnt_o and np.array_equal(labels[p], lab_o), f"case {case} pair {p} n={n}"


@pytest.mark.parametrize("block", range(3))
def test_production_engines_on_ragged_instances(block):
    """K2a (graph-coloured sparse) and K3 (dense tensor-core path) on the same ragged instances: no crash on n = 1,
    empty rows or isolated spins, states stay +-1, and the engines' energies are those of the states they return
    (fixed-point on K2a, bf16-split fields on K3: tolerances, not bit-exactness -- the returned fp64 energies of the
    drop-in classes come from K4)."""
    from nlmc_b200 import _lib, host
    from oracle import oracle as O
    rs = np.random.RandomState(3000 + block)
    for case in range(block * 6, block * 6 + 6):
        J, h, kind = random_instance(rs, case)
        n = J.shape[0]
        norm = max(np.max(np.abs(J)), 1e-30) if np.any(J) else 1.0
        J, h = J / norm, h / norm
        csr = O.Csr(J)
        prob = host.Problem(J, h)
        betas = np.linspace(0.3, 2.0, 5)
        col = _lib.Col(prob.inst, betas, seed=case)
        states, E = col.sweep_record(4)
        assert states.shape == (4, 5, n) and set(np.unique(states)) <= {-1, 1}
        Eo = O.energy(csr, h, states.reshape(-1, n)).reshape(4, 5)
        np.testing.assert_allclose(E, Eo, rtol=1e-6, atol=1e-6 * max(1, n), err_msg=f"K2a case {case} n={n}")
        col.close()
        d = _lib.Dense(prob.inst, betas, n_split=3, seed=case)
        d.sweep(3)
        S = d.get_spins()
        assert S.shape == (5, n) and set(np.unique(S)) <= {-1, 1}
        np.testing.assert_allclose(d.energies(), O.energy(csr, h, S), rtol=1e-4, atol=1e-4 * max(1, n),
                                   err_msg=f"K3 case {case} n={n}")
        d.close()


@pytest.mark.parametrize("kind", ["upper_int", "asym_int", "asym_real"])
def test_k1_asymmetric_j(kind):
    """The reference takes any J through J.dot(m) (NMC/nmc.py:86): an integer J that is not symmetric must not take
    the incremental-field kernel (which pushes J_kj into field j); the production engines refuse it."""
    from nlmc_b200 import _lib, host
    from oracle import oracle as O
    rs = np.random.RandomState({"upper_int": 1, "asym_int": 2, "asym_real": 3}[kind])
    n = 40
    U = rs.rand(n, n) < 0.3
    np.fill_diagonal(U, False)
    V = rs.randn(n, n) if kind == "asym_real" else rs.choice([-2.0, -1.0, 1.0, 3.0], size=(n, n))
    J = np.where(U, V, 0.0)
    if kind == "upper_int":
        J = np.triu(J)
    h = rs.randint(-1, 2, size=n).astype(float)
    csr = O.Csr(J)
    prob = host.Problem(J, h)
    assert _lib.lib().nlmc_instance_is_symmetric(prob.inst._h) == 0
    R, S = 2, 3
    reps = _lib.Replicas(prob.inst, R)
    m0 = rs.choice([-1, 1], size=(R, n)).astype(np.int8)
    sched = np.array([[0.8] * S, [1.7] * S])
    perm = np.stack([np.stack([rs.permutation(n) for _ in range(S)]) for _ in range(R)]).astype(np.int32)
    u = rs.rand(R, S, n)
    reps.set_spins(m0)
    M, E = reps.sweep_replay(perm, u, sched, prob.tanh_lut(sched) if prob.lut_half else None, prob.lut_half)
    for r in range(R):
        Mo, _ = O.mcmc(csr, h, m0[r], sched[r], perm=perm[r], u=u[r])
        assert np.array_equal(M[r], Mo)
        np.testing.assert_allclose(E[r], O.energy(csr, h, Mo), rtol=1e-9, atol=1e-12)
    for make in (lambda: _lib.Col(prob.inst, np.array([1.0]), seed=0), lambda: _lib.Dense(prob.inst, np.array([1.0]), seed=0)):
        with pytest.raises(_lib.NlmcError):
            make()
    # the symmetric part alone is accepted again
    Js = J + J.T
    assert _lib.lib().nlmc_instance_is_symmetric(host.Problem(Js, h).inst._h) == 1
