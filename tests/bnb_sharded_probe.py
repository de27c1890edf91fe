"""GPU box, torchrun: multi-GPU B&B node throughput (gomilp_b200/sharded.py over NCCL + gm_solve_wave).
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/bnb_sharded_probe.py"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import gomilp_b200 as gm
from gomilp_b200 import status as S
from gomilp_b200.sharded import gpu_wave_solver, milp_solve_sharded
from problems import knapsack

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
gm.init(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
for (n, m, lim) in [(30, 5, 8000), (60, 10, 8000)]:
    p = knapsack(np.random.default_rng(7), n, m)
    for rep in range(2):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = milp_solve_sharded(p["c"], None, None, p["G"], p["h"], p["integrality"], solve_wave=gpu_wave_solver(),
                               device=dev, mode=S.GM_BNB_FIXED, heuristic=S.GM_BRANCH_MOST_INFEASIBLE, node_limit=lim)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    if rank == 0:
        print(json.dumps({"knapsack": [n, m], "n_gpus": world, "status": r.status, "nodes": r.nodes, "waves": r.waves,
                          "pivots": r.pivots, "wall_s": dt, "nodes_per_s": r.nodes / dt, "device_ms": r.device_ms,
                          "exchange_bytes": r.exchange_bytes, "z": r.z}))
if world > 1:
    dist.destroy_process_group()
