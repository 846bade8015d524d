"""Multi-GPU layer of the production path: one process per GPU (torch.distributed).

Two partitions:
  * ShardedLadders     independent ladders in contiguous blocks of 128 over the ranks -- no exchange step at all;
  * ShardedBetaLadder  ONE set of ladders whose temperature range is cut into contiguous blocks, one per rank.  Per swap
                       round the only data that crosses the GPUs is one float64 energy per replica (all-gather over
                       NCCL / NVLink); every rank then takes the identical Philox-keyed exchange decisions and permutes
                       its copy of the beta labels (NPT/npt.py:649-680 in the label form of SURVEY D4).  Spin
                       configurations never move.

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


def beta_shard(n_beta: int, world: int, rank: int):
    """Contiguous block of temperature slots owned by `rank`: (first, count); the first n_beta % world ranks hold one
    slot more."""
    per, extra = divmod(n_beta, world)
    return rank * per + min(rank, extra), per + (1 if rank < extra else 0)


class ShardedBetaLadder:
    """`n_ladders` NPT ladders of one +-J instance with the temperature range sharded over the ranks of a process group.

    Per round: sweeps of the local slots -> bit-sliced energies into the all-gather send buffer -> all_gather (8 bytes per
    replica) -> identical label exchange on every rank -> local heat-bath thresholds rebuilt from the new labels.
    Everything is queued on one CUDA stream (the handle runs on the stream the collective is ordered on), so a round
    costs no host synchronisation.  Random streams are keyed by global slot and ladder indices: the ensemble evolves
    bit for bit like a single handle owning all slots (tests/test_gpu_label_exchange.py, tools/multigpu_beta_shard_check.py).
    """

    def __init__(self, prob, betas, n_ladders: int, seed: int, group=None, msc_factory=None, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.betas = np.ascontiguousarray(betas, dtype=np.float64)
        self.n_beta = len(self.betas)
        if self.n_beta < self.world:
            raise ValueError(f"{self.n_beta} temperatures cannot be sharded over {self.world} ranks")
        self.first, self.count = beta_shard(self.n_beta, self.world, self.rank)
        if msc_factory is None:
            from . import _lib
            msc_factory = lambda **kw: _lib.Msc(prob.inst, **kw)  # noqa: E731
        self.msc = msc_factory(betas=self.betas, n_ladders=n_ladders, seed=seed, labelled=True,
                               slot_begin=self.first, slot_count=self.count)
        self.n_ladders = self.msc.n_ladders
        on_gpu = device is not None or (dist.is_initialized() and dist.get_backend(group) == "nccl")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device) if on_gpu \
            else torch.device("cpu")
        self.stream = None
        if self.device.type == "cuda":
            self.stream = torch.cuda.Stream(self.device)
            self.msc.set_stream(self.stream.cuda_stream)
        self.max_count = beta_shard(self.n_beta, self.world, 0)[1]
        self.equal = self.n_beta % self.world == 0
        self.E_full = torch.zeros((self.n_beta, self.n_ladders), dtype=torch.float64, device=self.device)
        self.E_send = torch.zeros((self.max_count, self.n_ladders), dtype=torch.float64, device=self.device)
        self.E_recv = None if self.equal else torch.zeros((self.world, self.max_count, self.n_ladders),
                                                          dtype=torch.float64, device=self.device)

    def _on_stream(self):
        import contextlib
        return self.torch.cuda.stream(self.stream) if self.stream is not None else contextlib.nullcontext()

    def gather_energies(self):
        """Energies of the local slots -> E_full [n_beta][n_ladders] on every rank (slot-major)."""
        with self._on_stream():
            self.msc.energies_into(self.E_send[:self.count])
            if self.world == 1:
                self.E_full.copy_(self.E_send[:self.count])
            elif self.equal:
                self.dist.all_gather_into_tensor(self.E_full, self.E_send, group=self.group)
            else:
                self.dist.all_gather_into_tensor(self.E_recv.view(-1, self.n_ladders), self.E_send, group=self.group)
                for r in range(self.world):
                    f, c = beta_shard(self.n_beta, self.world, r)
                    self.E_full[f:f + c].copy_(self.E_recv[r, :c])
        return self.E_full

    def round(self, n_sweeps: int, num_swapping_pairs: int):
        with self._on_stream():
            self.msc.sweep(n_sweeps)
            self.gather_energies()
            self.msc.exchange_labels_from(self.E_full, num_swapping_pairs)

    def synchronize(self):
        if self.stream is not None:
            self.stream.synchronize()

    def labels(self) -> np.ndarray:
        """labels[slot][ladder] (identical on every rank)."""
        self.synchronize()
        return self.msc.labels()

    def energies_by_beta(self) -> np.ndarray:
        """[n_beta][n_ladders]: energy of the replica currently at temperature index i of each ladder."""
        E = self.gather_energies()
        self.synchronize()
        E = E.cpu().numpy()
        lab = self.labels().astype(np.int64)
        out = np.empty_like(E)
        np.put_along_axis(out, lab, E, axis=0)
        return out

    def close(self):
        self.synchronize()
        self.msc.close()
