"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed for plumbing.

The reference's only parallelism is data parallel over independent units with a
sum-reduction in the parent (examples/generate_opening_book.py:294-332,458-496;
benchmarks/agreement.py:72).  The same shape here: every rank runs the kernels on its
shard -- there is NO collective on the data path -- and ONE all-reduce of the integer
counters closes the job (NCCL over NVLink on GPUs, gloo in the CPU tests).

Because a playout is a pure function of (seed, root index, rollout index) the result is
bit-identical for any world size.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import numpy as np

from . import batch as B


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of `total` units for `rank`; sizes differ by at most one."""
    base, extra = divmod(int(total), int(world))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def _world(group=None) -> Tuple[int, int]:
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def all_reduce_counts(counts: np.ndarray, device: Optional[int] = None, group=None) -> np.ndarray:
    """Sum an integer counter array over all ranks (int64 on the wire)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return counts
    t = torch.from_numpy(np.ascontiguousarray(counts).astype(np.int64))
    if dist.get_backend(group) == "nccl":
        t = t.cuda(device if device is not None else torch.cuda.current_device())
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy().astype(counts.dtype)


def sharded_playout_tally(root, total_rollouts: int, seed: int = B.SEED_DEFAULT, max_plies: int = -1,
                          device: Optional[int] = None, group=None,
                          playouts_fn: Callable[..., Dict[str, np.ndarray]] = B.playouts) -> np.ndarray:
    """BASELINE config 1: `total_rollouts` playouts from ONE root, rollout-id ranges per rank,
    all-reduced [wins_p0, wins_p1, draws]."""
    rank, world = _world(group)
    begin, end = shard_range(total_rollouts, rank, world)
    local = np.zeros(3, np.int64)
    done = begin
    while done < end:   # rollouts_per_root is a uint32 in the ABI
        n = min(end - done, 1 << 31)
        r = playouts_fn(np.asarray(root, np.uint16).reshape(1, 8), n, seed=seed, max_plies=max_plies,
                        first_rollout=done, device=device)
        local += np.array([int(np.asarray(r["wins_p0"])[0]), int(np.asarray(r["wins_p1"])[0]),
                           int(np.asarray(r["draws"])[0])], np.int64)
        done += n
    return all_reduce_counts(local, device, group)


def sharded_root_tallies(roots: np.ndarray, rollouts: int, seed: int = B.SEED_DEFAULT, max_plies: int = -1,
                         device: Optional[int] = None, group=None,
                         playouts_fn: Callable[..., Dict[str, np.ndarray]] = B.playouts) -> np.ndarray:
    """BASELINE config 4: every rank evaluates a contiguous range of roots; the (n_roots, 3)
    tally matrix is assembled with one all-reduce (each rank contributes zeros elsewhere)."""
    rank, world = _world(group)
    roots = np.asarray(roots, np.uint16).reshape(-1, 8)
    begin, end = shard_range(len(roots), rank, world)
    out = np.zeros((len(roots), 3), np.int64)
    if end > begin:
        r = playouts_fn(roots[begin:end], rollouts, seed=seed, max_plies=max_plies, root_index_base=begin, device=device)
        out[begin:end, 0] = np.asarray(r["wins_p0"])
        out[begin:end, 1] = np.asarray(r["wins_p1"])
        out[begin:end, 2] = np.asarray(r["draws"])
    return all_reduce_counts(out, device, group)


def sharded_perft(roots: np.ndarray, depth: int, mult: Optional[np.ndarray] = None, device: Optional[int] = None,
                  group=None, perft_fn: Callable[..., Dict[str, np.ndarray]] = B.perft) -> Dict[str, np.ndarray]:
    """BASELINE configs 2/5: roots dealt round-robin (rank r takes roots r, r+world, ...: sorted
    canonical roots of similar size end up on different ranks), one all-reduce of the
    5 x (depth+1) counters."""
    rank, world = _world(group)
    roots = np.asarray(roots, np.uint16).reshape(-1, 8)
    mine = roots[rank::world]
    m = None if mult is None else np.asarray(mult, np.uint64)[rank::world]
    names = ["nodes", "wins_p0", "wins_p1", "blocked_p0_loses", "blocked_p1_loses"]
    local = np.zeros((5, depth + 1), np.int64)
    if len(mine):
        p = perft_fn(mine, depth, mult=m, device=device)
        for i, k in enumerate(names):
            local[i] = p[k].astype(np.int64)
    tot = all_reduce_counts(local, device, group)
    return {k: tot[i].astype(np.uint64) for i, k in enumerate(names)}
