"""Run under torchrun on N GPUs: ONE set of 128 ladders whose temperature range is sharded over the ranks
(distributed.ShardedBetaLadder: sweeps -> energies -> NCCL all-gather of 8 bytes per replica -> identical label
exchange on every rank) must equal the same ladders evolved in ONE labelled handle on rank 0 -- packed spins, labels
and energies bit for bit -- with exchanges crossing the rank boundaries.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/multigpu_beta_shard_check.py [L] [n_beta] [rounds]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "nonlocal-monte-carlo_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from nlmc_b200 import _lib, host, instances  # noqa: E402
from nlmc_b200.distributed import ShardedBetaLadder, beta_shard  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 12
n_beta = int(sys.argv[2]) if len(sys.argv) > 2 else 32
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 8
A, h = instances.ea3d_pm_j(L, 9)
prob = host.Problem(A, h, device=local)
betas = np.linspace(0.3, 1.8, n_beta)
ens = ShardedBetaLadder(prob, betas, 128, seed=77)
for _ in range(rounds):
    ens.round(4, max(1, n_beta // 3))
ens.synchronize()
labels = ens.labels()
E = ens.gather_energies().cpu().numpy()
local_packed = ens.msc.get_packed()
G = ens.n_ladders // 32
# every rank ships its block of packed spins to rank 0 for the comparison (a check, not part of the path)
parts = [None] * world
dist.all_gather_object(parts, (ens.first, ens.count, local_packed))
ok = True
if rank == 0:
    ref = _lib.Msc(prob.inst, betas, 128, seed=77, labelled=True)
    for _ in range(rounds):
        ref.round(4, max(1, n_beta // 3))
    P = ref.get_packed()
    same_spins = all(np.array_equal(p, P[:, f * G:(f + c) * G]) for f, c, p in parts)
    same_labels = bool(np.array_equal(ref.labels(), labels))
    same_E = bool(np.array_equal(ref.energies(), E))
    crossed = 0
    for r in range(1, world):
        f = beta_shard(n_beta, world, r)[0]
        crossed += int(np.sum(labels[f:] < f))          # temperatures that started below the cut and sit above it
    ok = same_spins and same_labels and same_E and crossed > 0
    print(json.dumps({"world": world, "L": L, "n_beta": n_beta, "rounds": rounds, "spins_equal": same_spins,
                      "labels_equal": same_labels, "energies_equal": same_E, "labels_across_rank_boundaries": crossed,
                      "accepted_per_round": ref.swap_counts(rounds).tolist(), "ok": ok}), flush=True)
    ref.close()
flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
dist.broadcast(flag, 0)
ens.close()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
