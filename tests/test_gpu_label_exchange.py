"""GPU: replica exchange by beta labels (north_star 4 / SURVEY D4) and the temperature range of a ladder sharded over
several handles -- the building block of the multi-GPU partition (distributed.ShardedBetaLadder).

  * a ladder split into blocks of slots, driven block by block with the energies gathered in between, evolves bit for bit
    like the single handle that owns every slot -- with exchanges crossing the block boundary;
  * the label form samples the exact Boltzmann distribution (full enumeration), energies regrouped by label;
  * short site rows (a few slots per block: several sites per warp) agree with the one-site-per-warp mapping."""
import itertools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nl():
    from nlmc_b200 import _lib, host
    return type("NL", (), dict(lib=_lib, host=host))


def _gathered_round(blocks, E_full, n_sweeps, pairs):
    """One round of a ladder sharded over `blocks` (handles on one GPU): what ShardedBetaLadder does with NCCL."""
    for m in blocks:
        m.sweep(n_sweeps)
    for m in blocks:
        m.energies_into(E_full[m.slot_begin:m.slot_begin + m.n_beta])
    import torch
    torch.cuda.synchronize()
    for m in blocks:
        m.sync()
    for m in blocks:
        m.exchange_labels_from(E_full, pairs)
    for m in blocks:
        m.sync()


@pytest.mark.parametrize("L,n_beta,cuts", [(6, 8, [0, 4, 8]), (6, 8, [0, 1, 3, 8]), (4, 32, [0, 16, 32]), (4, 32, [0, 8, 16, 24, 32])])
def test_sharded_slots_equal_single_handle(nl, L, n_beta, cuts):
    import torch
    from oracle import oracle as O
    A, h = O.ea3d_pm_j(L, 3)
    prob = nl.host.Problem(A, h)
    betas = np.linspace(0.3, 1.6, n_beta)
    whole = nl.lib.Msc(prob.inst, betas, 128, seed=77, labelled=True)
    blocks = [nl.lib.Msc(prob.inst, betas, 128, seed=77, labelled=True, slot_begin=a, slot_count=b - a)
              for a, b in zip(cuts[:-1], cuts[1:])]
    E_full = torch.zeros((n_beta, 128), dtype=torch.float64, device="cuda")
    G = whole.n_ladders // 32
    crossed = 0
    for rnd in range(12):
        whole.round(3, 3)
        _gathered_round(blocks, E_full, 3, 3)
        lab = whole.labels()
        for m in blocks:
            assert np.array_equal(m.labels(), lab), rnd
            P = m.get_packed()
            assert np.array_equal(P, whole.get_packed()[:, m.slot_begin * G:(m.slot_begin + m.n_beta) * G]), rnd
        assert np.array_equal(E_full.cpu().numpy(), whole.energies())
        # a label below the first cut sitting in a slot above it: an exchange crossed the block boundary
        crossed += int(np.sum(lab[cuts[1]:] < cuts[1]))
    assert crossed > 0
    assert np.array_equal(np.sort(lab, axis=0), np.arange(n_beta)[:, None] * np.ones((1, 128), dtype=int))  # a permutation
    assert np.array_equal(whole.swap_counts(12), blocks[0].swap_counts(12)) and whole.swap_counts(12).sum() > 0
    for m in blocks + [whole]:
        m.close()


def test_label_exchange_samples_exact_boltzmann(nl):
    """2D 4x4 periodic +-J, 4 temperatures, 1024 ladders: <E>(beta) regrouped by label against full enumeration, 3 sigma."""
    from test_gpu_production_msc import exact_mean_energy, lattice_2d
    A, h = lattice_2d(4, 11)
    betas = np.array([0.2, 0.5, 0.9, 1.4])
    exact = exact_mean_energy(A, betas)
    prob = nl.host.Problem(A, h)
    msc = nl.lib.Msc(prob.inst, betas, 1024, seed=4242, labelled=True)
    for _ in range(40):
        msc.round(5, 2)
    acc = np.zeros((len(betas), msc.n_ladders))
    T = 200
    for _ in range(T):
        msc.round(4, 2)
        E, lab = msc.energies(), msc.labels().astype(np.int64)
        by_beta = np.empty_like(E)
        np.put_along_axis(by_beta, lab, E, axis=0)
        acc += by_beta
    per_ladder = acc / T
    assert msc.swap_count() > 0 and not np.array_equal(lab, np.arange(4)[:, None] * np.ones((1, msc.n_ladders), dtype=int))
    for b in range(len(betas)):
        mean, err = per_ladder[b].mean(), per_ladder[b].std(ddof=1) / np.sqrt(per_ladder.shape[1])
        assert abs(mean - exact[b][0]) <= 3.0 * err + 1e-9, (betas[b], mean, exact[b][0], err)
    msc.close()


def test_label_and_bit_exchange_agree_without_swaps(nl, monkeypatch):
    """With no exchange the label form is the classic engine: identical trajectories (same streams, thresholds from the
    bit planes instead of three scalars), for one-site-per-warp rows (W = 128) and short rows (W = 8) -- at the same number
    of unconditional comparison steps (the label form defaults to 7, the scalar form to 6)."""
    from oracle import oracle as O
    monkeypatch.setenv("NLMC_MSC_STEPS", "6")
    A, h = O.ea3d_pm_j(6, 9)
    prob = nl.host.Problem(A, h)
    for n_beta in (32, 2):
        betas = np.linspace(0.2, 2.0, n_beta)
        a = nl.lib.Msc(prob.inst, betas, 128, seed=5)
        b = nl.lib.Msc(prob.inst, betas, 128, seed=5, labelled=True)
        a.sweep(7)
        b.sweep(7)
        assert np.array_equal(a.get_packed(), b.get_packed())
        assert np.array_equal(a.energies(), b.energies())
        a.close(), b.close()


def test_sharded_block_refuses_local_round(nl):
    from oracle import oracle as O
    A, h = O.ea3d_pm_j(4, 1)
    prob = nl.host.Problem(A, h)
    m = nl.lib.Msc(prob.inst, np.linspace(0.5, 1.0, 4), 128, seed=1, labelled=True, slot_begin=2, slot_count=2)
    with pytest.raises(nl.lib.NlmcError, match="gather"):
        m.round(1, 1)
        m.sync()
    with pytest.raises(nl.lib.NlmcError):
        m.set_betas(np.array([0.1, 0.2]))
    m.close()
