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
