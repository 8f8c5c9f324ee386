#!/usr/bin/env python3
"""The exchange step of the path at N GPUs: (1) the complete canonical game tree with the
frontier partitioned over the ranks by payload hash, one all-to-all of the locally merged child classes per depth;
(2) global canonical dedup of a random-image set (BASELINE config 3 scaled up).  Both are compared with the
single-GPU result computed on rank 0.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/multi_gpu_bfs.py [depth] [log2 states per rank]"""
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

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 16
log2n = int(sys.argv[2]) if len(sys.argv) > 2 else 22
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))


def timed(fn):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return out, float(dt.item())


D.sharded_symmetry_table(depth, device=local)                   # warm-up: context, memory pool, NCCL channels
rows, t_bfs = timed(lambda: D.sharded_symmetry_table(depth, device=local))
# the same with the exchange fused into the merge kernels (peer stores into the owners' inboxes)
boxes = D.PeerInboxes(1 << 14, local, copies=2)
D.fused_symmetry_table(6, device=local, inboxes=boxes)
D.fused_symmetry_table(depth, device=local, inboxes=boxes)      # sizes the inboxes for the deepest level
fused_rows, t_bfs_fused = timed(lambda: D.fused_symmetry_table(depth, device=local, inboxes=boxes))
del boxes

# config 3 scaled: every rank draws 2^log2n mid-game positions (its own index range of the root stream), and
# pushes each through a pseudo-random symmetry so that equal classes are spread over all ranks
n = 1 << log2n
roots = Q.random_roots(n, first=rank * n, min_plies=3, span=6, device=local)          # numpy (n, 8)
import itertools  # noqa: E402
table = np.array([Q.batch.encode_transform(d4, False, perm) for perm in itertools.permutations(range(4)) for d4 in range(8)],
                 np.uint16)                                                           # the 192 transform codes
codes = table[((np.arange(n, dtype=np.uint64) * np.uint64(2654435761) + np.uint64(rank)) % np.uint64(192)).astype(np.int64)]
images = Q.apply_symmetry(roots, codes, device=local)
dev_images = torch.from_numpy(np.ascontiguousarray(images).view(np.int16)).to(f"cuda:{local}")
(own, own_mult, n_global), t_dedup = timed(lambda: D.sharded_canon_dedup(dev_images, None, device=local))
mass = D.all_reduce_counts(np.array([int(own_mult.sum())], np.int64), local)
dbox = D.PeerInboxes(int(1.25 * n / world) + 4096, local, copies=1)
D.fused_canon_dedup(dev_images, None, device=local, inboxes=dbox)
(f_own, f_mult, f_global), t_dedup_fused = timed(lambda: D.fused_canon_dedup(dev_images, None, device=local, inboxes=dbox))
f_mass = D.all_reduce_counts(np.array([int(f_mult.sum())], np.int64), local)

# single-GPU check on rank 0: gather every rank's inputs
if world > 1:
    as64 = dev_images.view(torch.int64)                          # NCCL has no int16
    gathered = [torch.empty_like(as64) for _ in range(world)]
    dist.all_gather(gathered, as64)
    everything = torch.cat(gathered)
else:
    everything = dev_images
if rank == 0:
    one_rows = Q.symmetry_table(depth, device=local)
    u, um = Q.canon_dedup(everything, None, device=local)
    print(json.dumps({"n_gpus": world, "bfs_depth": depth, "bfs_seconds": t_bfs, "bfs_matches_single_gpu": rows == [list(map(int, r)) for r in one_rows],
                      "canonical_states_total": int(sum(r[1] for r in rows)),
                      "dedup_states": int(everything.shape[0]), "dedup_seconds": t_dedup,
                      "dedup_keys_per_s": int(everything.shape[0]) / t_dedup, "dedup_unique_global": n_global,
                      "dedup_matches_single_gpu": n_global == int(u.shape[0]) and int(mass[0]) == int(everything.shape[0]),
                      "fused_bfs_seconds": t_bfs_fused, "fused_bfs_matches_single_gpu": fused_rows == [list(map(int, r)) for r in one_rows],
                      "fused_dedup_seconds": t_dedup_fused, "fused_dedup_keys_per_s": int(everything.shape[0]) / t_dedup_fused,
                      "fused_dedup_matches_single_gpu": f_global == int(u.shape[0]) and int(f_mass[0]) == int(everything.shape[0])}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
