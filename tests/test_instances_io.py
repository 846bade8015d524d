"""Instance parsers: identical to the reference's own parsers on the shipped files (when
the reference is mounted) and a known-answer regression on a shipped ground state (golden copy, runs anywhere)."""
import glob
import importlib.util
import os

import numpy as np
import pytest

from conftest import golden
from nlmc_b200 import instances
from oracle import oracle as O
from oracle import ref_loader as rl


def test_chimera_known_ground_state(tmp_path):
    g = golden("known_answer_chimera128")
    path = tmp_path / "001.txt"
    path.write_text(str(g["instance_text"]))
    J, h = instances.read_chimera_droplet(str(path))
    assert J.shape == (128, 128) and h.shape == (128, 1)
    assert (abs(J - J.T) > 0).nnz == 0
    s = (2 * g["gs_bits"].astype(np.int64) - 1).astype(np.int8)
    E = O.energy(O.Csr(J), h, s)[0]
    assert abs(E - float(g["gs_energy"])) < 1e-5  # shipped value is rounded to 6 decimals
    # a single spin flip can only raise the energy of a ground state
    flips = np.tile(s, (128, 1))
    flips[np.arange(128), np.arange(128)] *= -1
    assert np.all(O.energy(O.Csr(J), h, flips) >= E - 1e-9)


def test_dict_semantics_and_options(tmp_path):
    p = tmp_path / "inst.txt"
    p.write_text("# comment\n\n0 1 2.5\n1 0 -1.0\n2 2 7\n1 2 0.0\n0 3 4\n")
    J, h = instances.read_instance(str(p), 0, "field")
    A = J.toarray()
    assert A[0, 1] == A[1, 0] == -1.0          # the later line wins, both directions
    assert A[0, 3] == A[3, 0] == 4.0 and J.nnz == 4   # explicit zeros dropped like csr_matrix(dense)
    assert h[2, 0] == 7.0 and J.shape == (4, 4)
    J2, h2 = instances.read_instance(str(p), 0, "skip", flip_sign=True)
    assert np.array_equal(J2.toarray(), -A) and not h2.any()
    with pytest.raises(ValueError):
        instances.read_instance(str(p), 0, "both")


@pytest.mark.skipif(not rl.available(), reason="/root/reference not mounted")
@pytest.mark.parametrize("script,func,ours,pattern", [
    ("wishart_example.py", "txt_to_A_wishart", "read_wishart", "wishart_small/wishart_planting_N_10_alpha_0.50/*inst_1*.txt"),
    ("DCL_example.py", "txt_to_A_DCL", "read_dcl", "DCL_instances/C8/00.txt"),
    ("chimera_example.py", "txt_to_A_droplet", "read_chimera_droplet", "Chimera_droplet_instances/chimera128_spinglass_power/00[12].txt"),
    ("contrived_wishart_example.py", "txt_to_A_wishart_contrived_tree", "read_contrived_wishart", "contrived_wishart_backbone/*/*inst_1.txt"),
])
def test_parsers_match_reference(script, func, ours, pattern):
    rl._install_matplotlib_stub()
    ex = os.path.join(rl.REFERENCE_ROOT, "NMC", "examples")
    files = sorted(f for f in glob.glob(os.path.join(ex, pattern)) if "gs_energ" not in f and "sol" not in f)[:3]
    if not files:
        pytest.skip("no shipped instance matches " + pattern)
    src = open(os.path.join(ex, script)).read().split("def main")[0].replace("from nmc import NMC", "")
    ns = {}
    exec(compile(src, script, "exec"), ns)
    for f in files:
        Jr, hr = ns[func](f)
        Jo, ho = getattr(instances, ours)(f, flip_sign=False)
        assert Jr.shape == Jo.shape and (Jr != Jo).nnz == 0
        assert np.array_equal(np.asarray(hr).reshape(-1), ho.reshape(-1))


def test_contrived_wishart_tree_generator_matches_reference(tmp_path):
    """The contrived 'Wishart backbone + trees' generator (contrived_instance_generator.py) reproduced draw for draw:
    adjacency, weights, cross connections, edge removal, fields and the written instance text."""
    from conftest import golden
    from nlmc_b200 import instances as I
    g = golden("contrived_generator")
    a = g["args"]
    assert np.array_equal(I.tree_backbone_adjacency(4, 1), g["adjacency_4_1"])
    np.random.seed(int(g["seed"]))
    J, h = I.contrived_wishart_tree(g["J_backbone"], int(a[0]), a[1], a[2], a[3], int(a[4]), a[5], int(a[6]))
    assert np.array_equal(J, g["J"]) and np.array_equal(h, g["h"])
    assert np.array_equal(J, J.T)
    path = tmp_path / "inst.txt"
    I.write_instance(J, h, str(path))
    assert path.read_text() == str(g["text"])
    # and back through the parser of the contrived example (sign convention: the file holds -J, -h)
    J2, h2 = I.read_contrived_wishart(str(path))
    np.testing.assert_allclose(J2.toarray(), J, rtol=0, atol=1e-15)
    np.testing.assert_allclose(h2.reshape(-1), h, rtol=0, atol=1e-15)


def test_synthetic_generators_match_the_oracle_definitions():
    """The product's generators of the benchmark instances and the oracle's are the same instances."""
    from nlmc_b200 import instances as I
    from oracle import oracle as O
    for L, seed in ((2, 1), (4, 3), (6, 5)):
        A, h = I.ea3d_pm_j(L, seed)
        B, _ = O.ea3d_pm_j(L, seed)
        assert (A != B).nnz == 0 and np.array_equal(A.indices, B.indices) and not h.any()
    J, _ = I.random_pm_graph(60, 0.1, 7)
    assert np.array_equal(J, O.random_pm_graph(60, 0.1, 7)[0])
    K, _ = I.sk_gaussian(30, 9)
    assert np.array_equal(K, O.sk_gaussian(30, 9)[0])
