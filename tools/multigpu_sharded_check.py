"""Run under torchrun on N GPUs: 128*N ladders sharded over the ranks (NCCL all_gather of the energies) must equal the same
ensemble evolved in ONE handle on rank 0, bit for bit (streams are keyed by the global ladder index).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/multigpu_sharded_check.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "nonlocal-monte-carlo_b200"))
import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from nlmc_b200 import _lib, host
from nlmc_b200.distributed import ShardedLadders
from bench import ea3d_csr

A = ea3d_csr(12, 9)
prob = host.Problem(A, np.zeros(A.shape[0]), device=local)
betas = np.linspace(0.3, 1.8, 8)
total = 128 * world
ens = ShardedLadders(prob, betas, total, seed=77)
for _ in range(5):
    ens.round(6, 2)
E = ens.energies()                      # [n_beta][total] on every rank
ok = True
if rank == 0:
    ref = _lib.Msc(prob.inst, betas, total, seed=77)
    for _ in range(5):
        ref.round(6, 2)
    ok = bool(np.array_equal(ref.energies(), E))
    print(f"world={world}: sharded ensemble == single handle: {ok}; E shape {E.shape}; mean E/N coldest {E[-1].mean() / A.shape[0]:.4f}", flush=True)
    ref.close()
flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
dist.broadcast(flag, 0)
ens.close()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
