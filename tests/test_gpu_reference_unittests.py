"""The reference's own unit tests (NMC/unittests/test_nmc.py, NPT/unittests/test_{npt,apt_ICM,apt_preprocessor}.py)
replayed against the drop-in classes on the GPU: same instance generator, same call arguments (positional where
the reference passes them positionally), same assertions, same side-effect files.  Plus edge cases of the API."""
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
EPS = np.finfo(float).eps


def generate_random_J_h(N, column_h):
    """NMC/unittests/test_nmc.py:9-17 / NPT/unittests/test_npt.py:8-16 (unseeded there; seeded here)."""
    h = np.random.randn(N, 1) if column_h else np.random.randn(N)
    iu = np.triu_indices(N, 1)
    J = np.zeros((N, N))
    J[iu] = np.random.randn(len(iu[0]))
    J += J.T
    return J, h


@pytest.fixture(autouse=True)
def _seed(tmp_cwd):
    np.random.seed(123)
    random.seed(123)


@pytest.mark.parametrize("mode", ["replay", "production"])
def test_nmc_unittest(mode):
    """NMC/unittests/test_nmc.py:25-59"""
    from nlmc_b200 import NMC
    J, h = generate_random_J_h(10, column_h=False)
    obj = NMC(J, h, mode=mode)
    assert np.array_equal(obj.J, J) and np.array_equal(obj.h, h.reshape(-1))
    M_overall, energy_overall, min_energy = obj.run(int(1e2), int(1e1), 2, 1, 1, 20, 3, 3, 0.01, 0.9, 0.9999999,
                                                    0.999999, 10, EPS, use_hash_table=False)
    assert isinstance(M_overall, np.ndarray)
    assert isinstance(energy_overall, (list, np.ndarray))
    assert isinstance(min_energy, (float, np.float64))
    assert M_overall.shape == (10, 60)


@pytest.mark.parametrize("mode", ["replay", "production"])
def test_npt_unittest(mode):
    """NPT/unittests/test_npt.py:24-84"""
    from nlmc_b200 import NPT
    N = 10
    J, h = generate_random_J_h(N, column_h=True)
    npt = NPT(J, h, mode=mode)
    assert np.array_equal(npt.J, J) and np.array_equal(npt.h, h.reshape(-1))
    beta_list = np.array([0.5, 1.0, 1.5, 2.0])
    num_replicas = 4
    M, Energy = npt.run(beta_list=beta_list, num_replicas=num_replicas, doNMC=[False] * 2 + [True] * 2,
                        num_sweeps_MCMC=int(1e2), num_sweeps_read=int(1e2), num_swap_attempts=int(1e1),
                        num_swapping_pairs=round(0.3 * num_replicas), num_cycles=10, full_update_frequency=1, M_skip=1,
                        temp_x=20, global_beta=1 / 0.366838 * 5, lambda_start=3, lambda_end=0.01,
                        lambda_reduction_factor=0.9, threshold_initial=0.9999999, threshold_cutoff=0.999999,
                        max_iterations=10, tolerance=EPS, use_hash_table=False, num_cores=1)
    assert M.shape == (N * num_replicas, int(1e2) // int(1e1))
    assert Energy.shape == (num_replicas,)


def test_apt_icm_unittest():
    """NPT/unittests/test_apt_ICM.py:24-43"""
    from nlmc_b200 import APT_ICM
    N = 10
    J, h = generate_random_J_h(N, column_h=True)
    apt = APT_ICM(J, h)
    assert np.array_equal(apt.J, J) and np.array_equal(apt.h, h)
    beta_list = np.array([0.5, 1.0, 1.5, 2.0])
    M, Energy = apt.run(beta_list, num_replicas=4, num_sweeps_MCMC=int(1e1), num_sweeps_read=int(1e1),
                        num_swap_attempts=int(1e0), num_swapping_pairs=1, use_hash_table=0, num_cores=8)
    assert M.shape == (N * 4, apt.num_sweeps_MCMC * 10) and Energy.shape == (4,)


@pytest.mark.parametrize("mode", ["replay"])
def test_apt_preprocessor_unittest(mode):
    """NPT/unittests/test_apt_preprocessor.py:32-72"""
    from nlmc_b200 import APT_preprocessor
    J, h = generate_random_J_h(10, column_h=True)
    apt = APT_preprocessor(J, h, mode=mode)
    beta, sigma = apt.run(num_sweeps_MCMC=100, num_sweeps_read=100, num_rng=10, beta_start=0.5, alpha=1.25,
                          sigma_E_val=1000, beta_max=4, use_hash_table=0, num_cores=2)
    assert isinstance(beta, list) and isinstance(sigma, list)
    assert os.path.exists("beta_list_python.npy") and os.path.exists("sigma_list_python.npy")
    assert os.path.isdir(os.path.join("Results", "data"))
    with pytest.raises(ValueError):
        APT_preprocessor(J, h).run(num_sweeps_MCMC=-100)


def test_run_twice_is_idempotent_and_does_not_touch_inputs():
    """run() rebinds self.J/self.h to normalised copies and never writes the caller's arrays."""
    from nlmc_b200 import NPT
    J, h = generate_random_J_h(12, column_h=False)
    J0, h0 = J.copy(), h.copy()
    obj = NPT(J, h)
    kw = dict(num_sweeps_MCMC=20, num_sweeps_read=10, num_swap_attempts=2, num_swapping_pairs=1)
    np.random.seed(5); random.seed(5)
    M1, E1 = obj.run(np.array([0.5, 1.0, 2.0]), 3, [False] * 3, **kw)
    assert np.array_equal(J, J0) and np.array_equal(h, h0)
    assert np.isclose(np.max(np.abs(obj.J)), 1.0)
    np.random.seed(5); random.seed(5)
    M2, E2 = obj.run(np.array([0.5, 1.0, 2.0]), 3, [False] * 3, **kw)
    assert np.array_equal(M1, M2) and np.allclose(E1, E2)


def test_error_contract():
    from nlmc_b200 import NPT
    J, h = generate_random_J_h(8, column_h=False)
    with pytest.raises(ValueError, match="length of doNMC"):
        NPT(J, h).run(np.array([0.5, 1.0]), 2, [False])
    with pytest.raises(ValueError, match="Cannot find non-overlapping pairs"):
        NPT(J, h).run(np.array([0.5, 1.0, 1.5]), 3, [False] * 3, num_sweeps_MCMC=4, num_sweeps_read=4,
                      num_swap_attempts=2, num_swapping_pairs=2)
    with pytest.raises(ValueError, match="LBP diverged at initial lambda"):
        from nlmc_b200 import NMC
        # weak clamping and only two iterations allowed: "iteration == max_iterations-1" at lambda_start (nmc.py:142-144)
        NMC(J, h).run(10, 5, 1, 1, 1, 20, 3, 0.05, 0.01, 0.9, 0.9999999, 0.999999, 2, EPS)


def test_edge_cases_vs_oracle():
    """isolated spins, explicit zero couplings, a single replica, one sweep, M_skip > 1, list-valued h."""
    from nlmc_b200 import NMC, _lib, host
    from oracle import oracle as O
    J = np.zeros((7, 7))
    J[0, 1] = J[1, 0] = 1.0
    J[2, 3] = J[3, 2] = -1.0     # spins 4, 5, 6 are isolated
    h = [0.0, 0.5, 0.0, -0.25, 0.75, 0.0, 0.0]
    prob = host.Problem(J, np.array(h))
    reps = _lib.Replicas(prob.inst, 1)
    rs = np.random.RandomState(3)
    m0 = rs.choice([-1, 1], size=(1, 7)).astype(np.int8)
    perm = rs.permutation(7).astype(np.int32)[None, None]
    u = rs.rand(1, 1, 7)
    reps.set_spins(m0)
    M, E = reps.sweep_replay(perm, u, np.array([[1.7]]), None, 0)
    Mo, _ = O.mcmc(O.Csr(J), np.array(h), m0[0], np.array([1.7]), perm=perm[0], u=u[0])
    assert np.array_equal(M[0], Mo)
    np.testing.assert_allclose(E[0], O.energy(O.Csr(J), np.array(h), Mo), rtol=1e-12)
    # M_skip = 2: half the columns, same energies as the stored states
    Jg, hg = generate_random_J_h(9, column_h=False)
    np.random.seed(9)
    Mo2, Eo2, _ = NMC(Jg, hg).run(20, 6, 2, 1, 2, 20, 3, 3, 0.01, 0.9, 0.9999999, 0.999999, 50, EPS)
    assert Mo2.shape == (9, 2 * 3 * 3) and len(Eo2) == 18
    norm = np.max(np.abs(Jg))
    np.testing.assert_allclose(O.energy(O.Csr(Jg / norm), hg / norm, Mo2.T.astype(np.int8)), Eo2, rtol=1e-9)


@pytest.mark.parametrize("lattice", [True, False])
@pytest.mark.parametrize("spm", [1, 3])
def test_apt_icm_production_mode(lattice, spm):
    """APT_ICM in production mode on both engines (bit-packed for +-J lattices, dense otherwise): the reference's
    return contract, states of +-1, and energies that are the minimum over the first columns of the returned M."""
    from nlmc_b200 import APT_ICM
    from oracle import oracle as O
    if lattice:
        A, h = O.ea3d_pm_j(4, 8)
        J = A.toarray()
    else:
        J, h = generate_random_J_h(20, column_h=False)
        J = J / np.max(np.abs(J)); h = h / 4
    n = J.shape[0]
    betas = np.array([0.4, 0.8, 1.2])
    obj = APT_ICM(J, h, mode="production")
    M, E = obj.run(betas, 3, num_sweeps_MCMC=4 * spm, num_sweeps_read=2 * spm, num_swap_attempts=4, num_swapping_pairs=1)
    assert M.shape == (n * 3, spm * 10) and E.shape == (3,) and np.all(np.abs(M) == 1)
    spr = (2 * spm) // 4
    csr = O.Csr(J)
    for r in range(3):
        if spr:
            Er = O.energy(csr, np.asarray(h).reshape(-1), M[r * n:(r + 1) * n, :spr].T.astype(np.int8))
            assert abs(E[r] - Er.min()) < 1e-3
