"""NPT(J, h, mode='production').run under torchrun (one rank per GPU, temperature range sharded): sanity of what the
ranks return -- M is +-1 in the reference's layout on the ranks asked for it, Energy is identical on every rank, equals the
minimum recorded energy of each temperature and the K4 energy of the returned last column.
    torchrun --nproc-per-node N tools/multigpu_api_check.py"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nonlocal-monte-carlo_b200"))
from nlmc_b200 import NPT, _lib, host, instances  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
dev = int(os.environ["LOCAL_RANK"])
L, R, spm, rounds = 16, 10, 6, 5   # 10 slots over the ranks: unequal blocks when world does not divide 10
A, h = instances.ea3d_pm_j(L, 3)
n = A.shape[0]
betas = np.linspace(0.3, 1.8, R)
out = {}
for m_ranks in (None, (0,)):
    np.random.seed(7)
    obj = NPT(A, h, mode="production", device=dev)
    obj.num_runs = 128
    obj.m_on_ranks = m_ranks
    M, E = obj.run(betas, R, [False] * R, num_sweeps_MCMC=spm * rounds, num_sweeps_read=spm * rounds, num_swap_attempts=rounds,
                   num_swapping_pairs=3)
    Et = torch.tensor(E, device=f"cuda:{dev}")
    gathered = [torch.zeros_like(Et) for _ in range(world)]
    dist.all_gather(gathered, Et)
    same_E = all(bool(torch.equal(g, gathered[0])) for g in gathered)
    ok_M = True
    if m_ranks is None or rank in m_ranks:
        ok_M = M is not None and M.shape == (R * n, spm) and bool(np.all(np.abs(M) == 1.0))
        prob = host.Problem(A / host.max_abs(A), h, device=dev)
        for r in range(R):
            Mr = M[r * n:(r + 1) * n, :]
            e_cols = prob.inst.energy_states(np.ascontiguousarray(Mr.T.astype(np.int8)))
            ok_M = ok_M and np.isclose(e_cols.min(), E[r]) and np.array_equal(e_cols, obj._EE1_list[r])
    else:
        ok_M = M is None
    cold_lower = bool(E[-1] < E[0] < 0)
    out[str(m_ranks)] = {"same_E_on_all_ranks": same_E, "M_ok": bool(ok_M), "colder_is_lower": cold_lower}
oks = torch.tensor([int(all(all(v.values()) for v in out.values()))], device=f"cuda:{dev}")
dist.all_reduce(oks, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"world": world, "all_ranks_ok": bool(oks.item()), "rank0": out}), flush=True)
dist.destroy_process_group()
sys.exit(0 if oks.item() else 1)
