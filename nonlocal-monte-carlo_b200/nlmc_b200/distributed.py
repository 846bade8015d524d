"""Multi-GPU layer of the production path: one process per GPU (torch.distributed), independent ladders
sharded in contiguous blocks of 128 over the ranks.

The path shards without a data-path collective: ladders (independent NPT runs of the same instance)
never interact, and every random stream is keyed by the GLOBAL ladder index, so an ensemble evolves
bit-identically on 1, 2, 4 or 8 GPUs.  The only communication is the gather of results (one float64
energy per replica) -- `all_gather` over NCCL on GPUs (gloo in the CPU tests).  A single instance is
never split across GPUs and spin configurations never leave their GPU (BASELINE.json north_star).
"""
from __future__ import annotations

import numpy as np

LANE_BLOCK = 128  # ladders per block: four 32-bit words sharing one beta (see csrc/nlmc_msc.cu)


def ladder_shard(n_ladders_total: int, world: int, rank: int):
    """Contiguous block of ladders owned by `rank`: (first, count), both multiples of 128 (count may be 0
    on trailing ranks when there are fewer blocks than ranks)."""
    blocks = (n_ladders_total + LANE_BLOCK - 1) // LANE_BLOCK
    per, extra = divmod(blocks, world)
    mine = per + (1 if rank < extra else 0)
    first = rank * per + min(rank, extra)
    return first * LANE_BLOCK, mine * LANE_BLOCK


class ShardedLadders:
    """`n_ladders_total` independent NPT ladders of one +-J instance over the ranks of a process group."""

    def __init__(self, prob, betas, n_ladders_total: int, seed: int, group=None, msc_factory=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.betas = np.asarray(betas, dtype=np.float64)
        self.n_total = ((n_ladders_total + LANE_BLOCK - 1) // LANE_BLOCK) * LANE_BLOCK
        self.first, self.count = ladder_shard(n_ladders_total, self.world, self.rank)
        if msc_factory is None:
            from . import _lib
            msc_factory = lambda **kw: _lib.Msc(prob.inst, **kw)  # noqa: E731
        self.msc = msc_factory(betas=self.betas, n_ladders=self.count, seed=seed,
                               ladder_offset=self.first) if self.count else None

    def round(self, n_sweeps: int, num_swapping_pairs: int):
        if self.msc is not None:
            self.msc.round(n_sweeps, num_swapping_pairs)

    def energies(self) -> np.ndarray:
        """[n_beta][n_ladders_total] on every rank, ladders in global order."""
        import torch
        local = self.msc.energies() if self.msc is not None else np.zeros((len(self.betas), 0))
        if self.world == 1:
            return local
        backend = self.dist.get_backend(self.group)
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        # equal-size buffers for all_gather: pad every shard to the largest one
        biggest = ladder_shard(self.n_total, self.world, 0)[1]
        buf = torch.zeros((len(self.betas), biggest), dtype=torch.float64, device=dev)
        buf[:, :local.shape[1]] = torch.from_numpy(np.ascontiguousarray(local)).to(dev)
        out = [torch.empty_like(buf) for _ in range(self.world)]
        self.dist.all_gather(out, buf, group=self.group)
        parts = []
        for r, t in enumerate(out):
            cnt = ladder_shard(self.n_total, self.world, r)[1]
            parts.append(t[:, :cnt].cpu().numpy())
        return np.concatenate(parts, axis=1)

    def close(self):
        if self.msc is not None:
            self.msc.close()
