"""Known-answer regressions on the B200: the production engines, driven as a parallel-tempering
search by tools/time_to_target.py, reach the planted ground-state energies that ship with the reference -- the Chimera
droplet instance through the dense tensor-core engine, the DCL instance through the graph-coloured sparse engine."""
import importlib.util
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ttt():
    spec = importlib.util.spec_from_file_location("time_to_target", os.path.join(ROOT, "tools", "time_to_target.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("name,engine", [("chimera128", "Dense"), ("dcl_c8", "Col"), ("wishart36", "Col")])
def test_production_engines_reach_planted_ground_state(ttt, name, engine, monkeypatch):
    if engine == "Dense":
        monkeypatch.setenv("NLMC_FORCE_DENSE", "1")  # the 128-spin Chimera instance would otherwise take K2a as well
    inst = ttt.load(name)
    seconds, sweeps, used = ttt.gpu_arm(inst, seed=3, runs=32)
    assert used == engine
    assert np.isfinite(seconds) and sweeps < ttt.MAX_ROUNDS * ttt.SPM


def test_known_answer_energy_conventions(ttt):
    """The fixtures' stated energies are attainable lower bounds under E = -(m^T J m/2 + m^T h) with J = -J_file."""
    from nlmc_b200 import host
    g = np.load(os.path.join(ROOT, "tests", "golden", "known_answer_chimera128.npz"))
    Jn, hn, target, _, _ = ttt.load("chimera128")
    prob = host.Problem(Jn, hn)
    s = (2 * g["gs_bits"].astype(np.int8) - 1)[None, :]
    E = prob.inst.energy_states(s)[0]          # K4 on the shipped ground-state bit string
    assert abs(E - target) < 1e-6


def test_reference_example_flow_on_chimera_droplet(ttt, tmp_cwd):
    """The flow of NPT/examples/chimera_example.py through the drop-in classes in production mode: parse the instance,
    adaptive ladder from APT_preprocessor, then NPT with NMC on the five coldest replicas -- the shipped ground-state
    energy is a lower bound that the cold replicas approach (within 10 % after these shortened runs; `Energy` is the
    minimum over the first sweeps of the last round only, npt.py:682-690) and the side-effect files appear."""
    import os
    import tempfile
    from nlmc_b200 import APT_preprocessor, NPT, instances
    g = np.load(os.path.join(ROOT, "tests", "golden", "known_answer_chimera128.npz"))
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
        f.write(str(g["instance_text"]))
    J, h = instances.read_chimera_droplet(f.name)          # J = -J_file, h = -h_file (chimera_example.py:49-50)
    os.unlink(f.name)
    np.random.seed(11)
    apt_prep = APT_preprocessor(J.copy(), h.copy(), mode="production")
    beta, sigma = apt_prep.run(num_sweeps_MCMC=300, num_sweeps_read=300, num_rng=64, beta_start=0.5, alpha=1.25,
                               sigma_E_val=1000, beta_max=64, use_hash_table=0, num_cores=8)
    beta_list = np.array(beta)
    assert 5 < len(beta_list) < 200 and np.all(np.diff(beta_list) > 0) and len(sigma) in (len(beta), len(beta) - 1)
    assert os.path.exists("beta_list_python.npy")
    R = len(beta_list)
    npt = NPT(J.toarray(), h, mode="production")
    M, Energy = npt.run(beta_list=beta_list, num_replicas=R, doNMC=[False] * (R - 5) + [True] * 5,
                        num_sweeps_MCMC=2000, num_sweeps_read=100, num_swap_attempts=10,
                        num_swapping_pairs=round(0.3 * R), num_cycles=10, full_update_frequency=1, M_skip=1, temp_x=20,
                        global_beta=1 / 0.366838 * 5, lambda_start=3, lambda_end=0.01, lambda_reduction_factor=0.9,
                        threshold_initial=0.9999999, threshold_cutoff=0.999999, max_iterations=100,
                        tolerance=np.finfo(float).eps, use_hash_table=False, num_cores=8)
    n = J.shape[0]
    assert M.shape == (R * n, 200) and Energy.shape == (R,)
    norm = abs(J).max()                                   # run() normalises by max |J| (npt.py:588-590)
    gs = float(g["gs_energy"]) / norm
    assert Energy.min() >= gs - 1e-6                      # nothing below the true ground state
    assert Energy.min() <= gs * 0.9                       # gs < 0: within 10 % of it


def test_preprocessor_production_matches_replay_on_real_couplings(ttt, tmp_cwd):
    """sigma_E(beta) is a property of the Boltzmann distribution: on an instance with real couplings and fields (generic
    production engine, not the bit-packed one) the production ladder follows the exact-replay ladder within the
    statistical error of 64 chains x 300 sweeps."""
    from nlmc_b200 import APT_preprocessor
    Jn, hn, _, _, _ = ttt.load("chimera128")
    out = {}
    for mode in ("replay", "production"):
        np.random.seed(5)
        obj = APT_preprocessor(Jn.copy(), hn.copy(), mode=mode)
        out[mode] = obj.run(num_sweeps_MCMC=300, num_sweeps_read=300, num_rng=64, beta_start=0.5, alpha=1.25,
                            sigma_E_val=1000, beta_max=3.0, use_hash_table=0, num_cores=1)
    (b_r, s_r), (b_p, s_p) = out["replay"], out["production"]
    k = min(len(s_r), len(s_p), 4)
    assert k >= 3
    np.testing.assert_allclose(s_p[:k], s_r[:k], rtol=0.15)
    np.testing.assert_allclose(b_p[:k], b_r[:k], rtol=0.15)


@pytest.mark.parametrize("mode", ["production", "replay"])
def test_nmc_example_flow_on_dcl(ttt, tmp_cwd, mode):
    """NMC/examples/DCL_example.py through the drop-in NMC (sweeps reduced): dense J from the parser, positional
    arguments as in the example, energies consistent with the returned states and bounded by the planted minimum."""
    import os
    import tempfile
    from nlmc_b200 import NMC, instances
    from oracle import oracle as O
    g = np.load(os.path.join(ROOT, "tests", "golden", "known_answer_dcl_c8.npz"))
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
        f.write(str(g["instance_text"]))
    J, h = instances.read_dcl(f.name)                     # J = -J_file (DCL_example.py:52-53)
    os.unlink(f.name)
    np.random.seed(3)
    sweeps = 300 if mode == "production" else 60
    nmc = NMC(J.toarray(), h, mode=mode)
    M, E, min_energy = nmc.run(sweeps, sweeps, 2, 1, 1, 20, 3, 3, 0.01, 0.9, 0.9999999, 0.999999, 100,
                               np.finfo(float).eps, use_hash_table=False)
    n = J.shape[0]
    assert M.shape == (n, 2 * 3 * sweeps) and E.shape == (2 * 3 * sweeps,) and min_energy == E.min()
    csr = O.Csr(J)
    cols = [0, len(E) // 2, len(E) - 1]
    np.testing.assert_allclose(E[cols], O.energy(csr, np.zeros(n), M[:, cols].T.astype(np.int8)), rtol=1e-9)
    target = float(g["min_energy"])
    assert min_energy >= target - 0.01                    # the file rounds 1/7: its optimum is 0.00175 below the stated one
    if mode == "production":
        assert min_energy <= 0.9 * target
