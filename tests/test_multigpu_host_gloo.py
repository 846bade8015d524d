"""N > 1 host path on CPU: world_size-2 `gloo` process group.  The bit-packed device state is replaced by
a deterministic stand-in whose "energies" depend only on (seed, beta, GLOBAL ladder index, round) -- the
property the real kernels have -- so the test pins the sharding arithmetic, the gather order and the
1-rank == 2-rank invariance of the distributed layer."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class StandInMsc:
    def __init__(self, betas, n_ladders, seed, ladder_offset):
        self.betas, self.n_ladders, self.seed, self.offset = betas, n_ladders, seed, ladder_offset
        self.rounds = 0

    def round(self, n_sweeps, pairs):
        self.rounds += n_sweeps

    def energies(self):
        lad = np.arange(self.offset, self.offset + self.n_ladders)
        return -(self.betas[:, None] * 1000 + lad[None, :] + 0.001 * self.rounds + self.seed)

    def close(self):
        pass


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "nonlocal-monte-carlo_b200"))
    from nlmc_b200.distributed import ShardedLadders
    ens = ShardedLadders(None, np.array([0.5, 1.0, 1.5]), total, seed=7, msc_factory=lambda **kw: StandInMsc(**kw))
    ens.round(4, 1)
    E = ens.energies()
    q.put((rank, ens.first, ens.count, E))
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("total", [256, 384, 128])
def test_two_rank_gather_matches_single_rank(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    single = StandInMsc(np.array([0.5, 1.0, 1.5]), total, 7, 0)
    single.round(4, 1)
    expect = single.energies()
    covered = []
    for rank, first, count, E in res:
        assert E.shape == expect.shape and np.array_equal(E, expect)  # every rank holds the full result
        assert first % 128 == 0 and count % 128 == 0
        covered += list(range(first, first + count))
    assert covered == list(range(total))


def test_ladder_shard_covers_without_overlap():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "nonlocal-monte-carlo_b200"))
    from nlmc_b200.distributed import ladder_shard
    for total in (1, 128, 129, 1024, 4096, 5000):
        for world in (1, 2, 3, 4, 8):
            spans = [ladder_shard(total, world, r) for r in range(world)]
            pos = 0
            for first, count in spans:
                assert first == pos and count % 128 == 0
                pos += count
            assert pos == ((total + 127) // 128) * 128
            counts = [c for _, c in spans]
            assert max(counts) - min(counts) <= 128


# ---------------------------------------------------------------------------------------------------
# temperature range of ONE ladder set sharded over the ranks: energy all-gather + identical label decisions
class StandInLabelled:
    """Stand-in for a labelled Msc block: energies depend only on (GLOBAL slot, ladder, round, current label), the
    exchange rule is the library's (adjacent temperatures, Metropolis on the gathered energies) with a counter-based
    stream keyed by (seed, round, ladder) -- so every rank must arrive at the same permutation."""

    def __init__(self, betas, n_ladders, seed, labelled, slot_begin, slot_count):
        assert labelled
        self.betas, self.n_ladders, self.seed = np.asarray(betas), ((n_ladders + 127) // 128) * 128, seed
        self.n_beta_total, self.slot_begin, self.n_beta = len(betas), slot_begin, slot_count
        self.lab = np.tile(np.arange(self.n_beta_total)[:, None], (1, self.n_ladders))
        self.round_no, self.sweeps = 0, 0

    def sweep(self, n):
        self.sweeps += n

    def energies_into(self, t):
        s = np.arange(self.slot_begin, self.slot_begin + self.n_beta)[:, None]
        l = np.arange(self.n_ladders)[None, :]
        lab = self.lab[self.slot_begin:self.slot_begin + self.n_beta]
        E = -100.0 * self.betas[lab] + ((s * 7919 + l * 104729 + self.sweeps * 31) % 97) - 48.0
        t.copy_(torch.from_numpy(E))

    def exchange_labels_from(self, E_full, pairs):
        E = E_full.numpy()
        for l in range(self.n_ladders):
            rs = np.random.RandomState((self.seed * 1000003 + self.round_no * 8191 + l) % (2 ** 31))
            avail = list(range(self.n_beta_total - 1))
            for _ in range(pairs):
                if not avail:
                    break
                i = avail[rs.randint(len(avail))]
                avail = [j for j in avail if abs(j - i) > 1]
                sa = int(np.where(self.lab[:, l] == i)[0][0]); sb = int(np.where(self.lab[:, l] == i + 1)[0][0])
                x = (self.betas[i + 1] - self.betas[i]) * (E[sb, l] - E[sa, l])
                if rs.rand() < min(1.0, np.exp(x)):
                    self.lab[sa, l], self.lab[sb, l] = i + 1, i
        self.round_no += 1

    def labels(self):
        return self.lab.astype(np.uint8)

    def close(self):
        pass


def _beta_worker(rank, world, port, n_beta, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "nonlocal-monte-carlo_b200"))
    from nlmc_b200.distributed import ShardedBetaLadder
    ens = ShardedBetaLadder(None, np.linspace(0.2, 1.5, n_beta), 128, seed=5, msc_factory=lambda **kw: StandInLabelled(**kw))
    for _ in range(6):
        ens.round(3, max(1, n_beta // 3))
    q.put((rank, ens.first, ens.count, ens.labels(), ens.energies_by_beta()))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_beta", [8, 7])
def test_beta_range_sharded_over_two_ranks_matches_single_rank(n_beta):
    """world_size 2 over gloo: equal (8) and unequal (7 = 4 + 3) blocks of slots."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_beta_worker, args=(r, 2, port, n_beta, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single rank: one stand-in owning every slot, driven by the same class without a process group
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "nonlocal-monte-carlo_b200"))
    from nlmc_b200.distributed import ShardedBetaLadder, beta_shard
    one = ShardedBetaLadder(None, np.linspace(0.2, 1.5, n_beta), 128, seed=5, msc_factory=lambda **kw: StandInLabelled(**kw))
    for _ in range(6):
        one.round(3, max(1, n_beta // 3))
    lab1, E1 = one.labels(), one.energies_by_beta()
    assert [(f, c) for _, f, c, _, _ in res] == [beta_shard(n_beta, 2, r) for r in range(2)]
    for _, _, _, lab, E in res:
        assert np.array_equal(lab, lab1) and np.array_equal(E, E1)
    cut = res[1][1]
    assert np.sum(lab1[cut:] < cut) > 0            # exchanges crossed the rank boundary
    assert np.array_equal(np.sort(lab1, axis=0), np.tile(np.arange(n_beta)[:, None], (1, 128)))


def test_beta_shard_covers_without_overlap():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "nonlocal-monte-carlo_b200"))
    from nlmc_b200.distributed import beta_shard
    for n_beta in (1, 2, 7, 32, 33, 128):
        for world in (1, 2, 3, 4, 8):
            if world > n_beta:
                continue
            pos = 0
            for r in range(world):
                f, c = beta_shard(n_beta, world, r)
                assert f == pos and c >= 1
                pos += c
            assert pos == n_beta
