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
    def probe_fn(roots, device=None):       # qk_perft_probe: move sequences of length 2 below every root
        return np.array([int(O.perft(r[None], 2)["nodes"][2]) for r in np.asarray(roots)], np.uint32)

    pf = D.sharded_perft(roots[:9], 2, mult=np.arange(1, 10, dtype=np.uint64), perft_fn=perft_fn, probe_fn=probe_fn)

    # the exchange step: global dedup and the partitioned BFS (oracle stand-ins for the two kernels)
    def dedup_fn(states, mult=None, color_swap=False, cap=None, device=None, sort_output=True):
        s = np.asarray(states, np.uint16).reshape(-1, 8)
        m = np.ones(len(s), np.uint64) if mult is None else np.asarray(mult, np.uint64)
        canon = O.canonical(s, color_swap)
        uniq, inv = np.unique(np.ascontiguousarray(canon).view("V16").ravel(), return_inverse=True)
        sums = np.zeros(len(uniq), np.uint64)
        np.add.at(sums, inv, m)
        return uniq.view(np.uint16).reshape(-1, 8), sums

    def bfs_fn(frontier, mult, cap=0, sort_output=False, device=None):
        f = np.asarray(frontier, np.uint16).reshape(-1, 8)
        m = np.asarray(mult, np.uint64)
        mask, tm, _ = O.movegen_masks(f)
        idx = [i for i, mk in enumerate(mask) for a in range(64) if (int(mk) >> a) & 1]
        act = [a for mk in mask for a in range(64) if (int(mk) >> a) & 1]
        kids = O.apply_moves(f[idx], np.array(act, np.uint8))
        km = m[idx]
        win = O.win_status(kids)
        st = {"total_moves": int(km.sum()), "line_wins_p0": int(km[win == 1].sum()), "line_wins_p1": int(km[win == 2].sum()),
              "blocked_p0": int(m[(mask == 0) & (tm == 0)].sum()), "blocked_p1": int(m[(mask == 0) & (tm == 1)].sum())}
        u, um = dedup_fn(kids[win == 0], km[win == 0])
        return u, um, st

    images = np.load(os.path.join(ROOT, "tests", "golden", "symmetry_images.npz"))
    pool = np.concatenate([images["image"], images["states"]]).astype(np.uint16)
    b, e = D.shard_range(len(pool), rank, world)
    own, own_mult, n_global = D.sharded_canon_dedup(pool[b:e], None, dedup_fn=dedup_fn)
    rows = D.sharded_symmetry_table(4, on_device=False, bfs_fn=bfs_fn, dedup_fn=dedup_fn)
    q.put(("shard", rank, np.asarray(own).tolist(), np.asarray(own_mult).tolist(), n_global, rows))
    if rank == 0:
        q.put(("main", tally.tolist(), per_root.tolist(), {k: v.tolist() for k, v in pf.items()}))
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
    got = []
    for _ in range(3):
        for _ in range(120):                    # a rank that died must fail the test at once, not after the timeout
            try:
                got.append(q.get(timeout=1))
                break
            except Exception:
                assert all(p.exitcode in (None, 0) for p in procs), [p.exitcode for p in procs]
        else:
            raise AssertionError("world-2 workers did not answer within 120 s")
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    _, tally, per_root, pf = next(g for g in got if g[0] == "main")
    shards = sorted((g for g in got if g[0] == "shard"), key=lambda g: g[1])
    one = O.playouts(np.zeros((1, 8), np.uint16), 5001)
    assert tally == [int(one["wins_p0"][0]), int(one["wins_p1"][0]), int(one["draws"][0])]
    roots = O.make_roots(37)
    r = O.playouts(roots, 64)
    assert np.array_equal(np.array(per_root), np.stack([r["wins_p0"], r["wins_p1"], r["draws"]], 1).astype(np.int64))
    want = O.perft(roots[:9], 2, np.arange(1, 10, dtype=np.uint64))
    for k in want:
        assert pf[k] == want[k].tolist()

    # global dedup: the owners' shares are disjoint, their union equals the single-process result
    images = np.load(os.path.join(ROOT, "tests", "golden", "symmetry_images.npz"))
    pool = np.concatenate([images["image"], images["states"]]).astype(np.uint16)
    canon = O.canonical(pool)
    uniq, counts = np.unique(np.ascontiguousarray(canon).view("V16").ravel(), return_counts=True)
    want = {bytes(u): int(c) for u, c in zip(uniq, counts)}
    have = {}
    for _, _, own, own_mult, n_global, _ in shards:
        assert n_global == len(want)
        for p_, m_ in zip(own, own_mult):
            key = np.array(p_, np.uint16).tobytes()
            assert key not in have
            have[key] = int(m_)
    assert have == want
    assert min(len(s[2]) for s in shards) > 0            # both ranks own something
    # partitioned BFS: the reference's table rows (GAME_TREE_ANALYSIS.md:4-7), identical on every rank
    table = O.symmetry_table(4)
    for s in shards:
        assert np.array_equal(np.array(s[5], np.uint64), table)
