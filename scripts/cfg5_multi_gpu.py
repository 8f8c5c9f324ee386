#!/usr/bin/env python3
"""BASELINE config 5: full-depth game-tree enumeration sharded by canonical depth-4 subtrees across the GPUs of one
box, one NCCL all-reduce of the node/win counters at the end.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/cfg5_multi_gpu.py [depth]"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantik-core-py_b200"))
import quantik_b200 as Q  # noqa: E402
from quantik_b200 import dist as D  # noqa: E402

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 12
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
d4 = np.load(os.path.join(ROOT, "tests", "golden", "depth4_canonical.npz"))
Q.perft(d4["states"][:64], 2, d4["mult"][:64], device=local)          # warm-up (context, pool)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
p = D.sharded_perft(d4["states"], depth, d4["mult"], device=local)
torch.cuda.synchronize()
dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{local}")
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
if rank == 0:
    nodes = int(p["nodes"][1:].astype(object).sum())
    rows = Q.symmetry_table(4 + depth, device=local)
    ok = True
    for r in range(1, depth + 1):
        want = rows[4 + r - 1]
        got = [int(p["nodes"][r]), int(p["wins_p0"][r]) + int(p["blocked_p1_loses"][r - 1]),
               int(p["wins_p1"][r]) + int(p["blocked_p0_loses"][r - 1]), int(p["nodes"][r]) - int(p["wins_p0"][r]) - int(p["wins_p1"][r])]
        ok &= got == [want[0], want[2], want[3], want[4]]
    print(json.dumps({"config": "cfg5: 10946 canonical depth-4 roots (weighted) to depth %d" % (4 + depth), "n_gpus": world,
                      "seconds": float(dt.item()), "weighted_nodes": nodes, "nodes_per_s": nodes / float(dt.item()),
                      "matches_canonical_bfs_every_depth": bool(ok)}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
