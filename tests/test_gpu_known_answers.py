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


@pytest.mark.parametrize("name,engine", [("chimera128", "Dense"), ("dcl_c8", "Col")])
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
