"""Pins the CPU oracle against the LIVE reference (imported from /root/reference with a matplotlib
stub).  Only runs where the reference is mounted (the build container); skipped on the GPU box."""
import warnings

import numpy as np
import pytest

from oracle import oracle as O
from oracle import ref_loader as rl

pytestmark = pytest.mark.skipif(not rl.available(), reason="/root/reference not mounted")
EPS = np.finfo(float).eps
warnings.simplefilter("ignore")


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_mcmc_random_instances(seed):
    rs = np.random.RandomState(seed)
    N = int(rs.randint(8, 40))
    J, h = O.random_pm_graph(N, 0.3, seed)
    if seed == 2:
        h = rs.choice([-1.0, 0.0, 1.0], size=N)
    obj = rl.nmc().NMC(J, h)
    rl.seed_all(seed)
    m0 = np.sign(2 * np.random.rand(N) - 1)
    st = np.random.get_state()
    Mref = obj.MCMC(7, m0.copy(), 1.9, J, h, anneal=(seed == 3))
    np.random.set_state(st)
    M, _ = O.mcmc(O.Csr(J), h, m0, O.anneal_schedule(7, 1.9, seed == 3, 1, 0))
    assert np.array_equal(M.T.astype(float), Mref)


def test_npt_run_fork_semantics():
    """num_cores=1: one worker forked at the first submit."""
    A, h = O.ea3d_pm_j(3, 9)
    J = A.toarray()
    betas = np.array([0.3, 0.9, 1.5])
    kw = dict(num_sweeps_MCMC=20, num_sweeps_read=10, num_swap_attempts=5, num_swapping_pairs=1)
    rl.seed_all(5)
    with rl.quiet_tmp_cwd():
        Mr, Er = rl.npt().NPT(J, h).run(betas, 3, [False] * 3, num_cores=1, **kw)
    rl.seed_all(5)
    Mo, Eo = O.npt_run(J, h, betas, 3, [False] * 3, **kw)
    assert np.array_equal(Mr, Mo) and np.array_equal(Er, Eo)


def test_reference_unit_test_contract():
    """The reference's own NPT unit test configuration (NPT/unittests/test_npt.py:30-84) through the oracle."""
    rs = np.random.RandomState(0)
    N = 10
    h = rs.randn(N, 1)
    J = np.zeros((N, N))
    iu = np.triu_indices(N, 1)
    J[iu] = rs.randn(len(iu[0]))
    J += J.T
    betas = np.array([0.5, 1.0, 1.5, 2.0])
    kw = dict(num_sweeps_MCMC=100, num_sweeps_read=100, num_swap_attempts=10, num_swapping_pairs=1, num_cycles=10,
              full_update_frequency=1, M_skip=1, temp_x=20, global_beta=1 / 0.366838 * 5, lambda_start=3,
              lambda_end=0.01, lambda_reduction_factor=0.9, threshold_initial=0.9999999, threshold_cutoff=0.999999,
              max_iterations=10, tolerance=EPS)
    rl.seed_all(8)
    with rl.quiet_tmp_cwd():
        Mr, Er = rl.npt().NPT(J, h).run(betas, 4, [False, False, True, True], num_cores=1, **kw)
    rl.seed_all(8)
    Mo, Eo = O.npt_run(J, h, betas, 4, [False, False, True, True], **kw)
    assert Mr.shape == (N * 4, 10)
    assert np.array_equal(Mr, Mo)
    np.testing.assert_allclose(Er, Eo, rtol=1e-9)
