"""world_size-2 gloo test of the N>1 path: shard ranges, root-index bases and the single
all-reduce of integer counters (quantik_b200.dist).  The oracle stands in for the kernels
(CPU box); the GPU tier runs the same functions with the CUDA path."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "quantik-core-py_b200"))
    import torch.distributed as dist
    import oracle as O
    from quantik_b200 import dist as D
    O.set_num_threads(2)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)

    def playouts_fn(roots, rollouts, seed, max_plies=-1, first_rollout=0, root_index_base=0, device=None):
        return O.playouts(roots, rollouts, seed, first_rollout, root_index_base, max_plies)

    def perft_fn(roots, depth, mult=None, device=None):
        return O.perft(roots, depth, mult)

    empty = np.zeros(8, np.uint16)
    tally = D.sharded_playout_tally(empty, 5001, seed=O.SEED_DEFAULT, playouts_fn=playouts_fn)
    roots = O.make_roots(37)
    per_root = D.sharded_root_tallies(roots, 64, seed=O.SEED_DEFAULT, playouts_fn=playouts_fn)
    pf = D.sharded_perft(roots[:9], 2, mult=np.arange(1, 10, dtype=np.uint64), perft_fn=perft_fn)
    if rank == 0:
        q.put((tally.tolist(), per_root.tolist(), {k: v.tolist() for k, v in pf.items()}))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions():
    from quantik_b200.dist import shard_range
    for total in (0, 1, 7, 8, 100001):
        for world in (1, 2, 3, 8):
            parts = [shard_range(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == total
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in parts]
            assert max(sizes) - min(sizes) <= 1


def test_world2_matches_single_process():
    import oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    tally, per_root, pf = q.get(timeout=180)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    one = O.playouts(np.zeros((1, 8), np.uint16), 5001)
    assert tally == [int(one["wins_p0"][0]), int(one["wins_p1"][0]), int(one["draws"][0])]
    roots = O.make_roots(37)
    r = O.playouts(roots, 64)
    assert np.array_equal(np.array(per_root), np.stack([r["wins_p0"], r["wins_p1"], r["draws"]], 1).astype(np.int64))
    want = O.perft(roots[:9], 2, np.arange(1, 10, dtype=np.uint64))
    for k in want:
        assert pf[k] == want[k].tolist()
