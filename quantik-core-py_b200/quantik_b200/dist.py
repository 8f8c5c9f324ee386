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


# ----------------------------------------------------------------------------- global dedup: the one exchange step
# Canonical keys shard by index range, but GLOBAL unique counts need equal payloads to meet on one rank
# (SURVEY 8e): every payload has an owner rank = hash(payload) mod world; one all-to-all moves each locally
# merged (payload, multiplicity) pair to its owner, which merges again.  torch is plumbing here: the hash,
# the bucket sort and the collective; canonicalisation and both merges are the library's kernels.
def _as_i64_pairs(payloads):
    """(U, 8) uint16 numpy / int16 torch -> torch int64 (U, 2) on the same device, no copy where possible."""
    import torch
    if isinstance(payloads, torch.Tensor):
        return payloads.contiguous().view(torch.int64).reshape(-1, 2)
    a = np.ascontiguousarray(payloads, dtype=np.uint16).reshape(-1, 8)
    return torch.from_numpy(a.view(np.int64).reshape(-1, 2))


def owner_rank(payloads, world: int):
    """Owner of each canonical payload: a 64-bit mix of its two halves, mod world (torch int64 tensor)."""
    import torch
    p = _as_i64_pairs(payloads)
    h = p[:, 0] * -7046029254386353131 + p[:, 1] * -4417276706812531889      # wraps mod 2^64
    h = h ^ (h >> 29)
    h = h * -4658895280553007687
    h = h ^ (h >> 32)
    return torch.remainder(h & 0x7FFFFFFFFFFFFFFF, world)


def exchange_by_owner(payloads, mult, group=None):
    """Send every (payload, multiplicity) pair to its owner rank with one all-to-all (three
    all_to_all_single calls: counts, payloads, multiplicities).  -> (received int64 (R, 2), received int64 (R,))
    on the device the inputs live on.  With one rank it is the identity."""
    import torch
    import torch.distributed as dist
    p = _as_i64_pairs(payloads)
    m = mult if isinstance(mult, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(mult).astype(np.int64))
    m = m.to(p.device).reshape(-1)
    rank, world = _world(group)
    if world == 1:
        return p, m
    if dist.get_backend(group) == "nccl" and not p.is_cuda:
        p, m = p.cuda(), m.cuda()
    owner = owner_rank(p, world)
    order = torch.argsort(owner, stable=True)
    send_p, send_m = p[order].contiguous(), m[order].contiguous()
    send_counts = torch.bincount(owner, minlength=world)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    ins, outs = send_counts.tolist(), recv_counts.tolist()
    recv_p = torch.empty((sum(outs), 2), dtype=torch.int64, device=p.device)
    recv_m = torch.empty((sum(outs),), dtype=torch.int64, device=p.device)
    dist.all_to_all_single(recv_p, send_p, outs, ins, group=group)
    dist.all_to_all_single(recv_m, send_m, outs, ins, group=group)
    return recv_p, recv_m


def _like_states(p64, host: bool):
    """int64 (U, 2) tensor -> what the batch API takes: numpy (U, 8) uint16 on the host, the tensor itself on the GPU."""
    if host or not p64.is_cuda:
        return p64.cpu().numpy().view(np.uint16).reshape(-1, 8)
    return p64


def sharded_canon_dedup(states, mult=None, color_swap: bool = False, device: Optional[int] = None, group=None,
                        dedup_fn: Callable[..., Tuple] = B.canon_dedup, sort_output: bool = False):
    """BASELINE config 3 at N GPUs: every rank passes ITS slice of the states; returns
    (payloads this rank owns, their global multiplicities, global number of distinct classes).
    local canonicalise+merge -> all-to-all by owner -> merge at the owner -> all-reduce of the counts."""
    import torch
    n = len(states)
    _, world = _world(group)
    if n:       # the local merge feeds the exchange: its order is irrelevant unless it is also the result
        uniq, um = dedup_fn(states, mult, color_swap=color_swap, device=device, sort_output=sort_output and world == 1)
    else:
        uniq, um = np.zeros((0, 8), np.uint16), np.zeros(0, np.uint64)
    host = not (isinstance(uniq, torch.Tensor) and uniq.is_cuda)
    rp, rm = exchange_by_owner(uniq, um, group)
    if world > 1 and len(rp):
        mine, mm = dedup_fn(_like_states(rp, host), rm if not host else rm.cpu().numpy().astype(np.uint64),
                            color_swap=color_swap, device=device, sort_output=sort_output)
    elif world > 1:
        mine, mm = np.zeros((0, 8), np.uint16), np.zeros(0, np.uint64)
    else:
        mine, mm = uniq, um
    total = all_reduce_counts(np.array([len(mine)], np.int64), device, group)
    return mine, mm, int(total[0])


def sharded_symmetry_table(max_depth: int, device: Optional[int] = None, group=None, on_device: bool = True,
                           bfs_fn: Callable[..., Tuple] = B.bfs_level, dedup_fn: Callable[..., Tuple] = B.canon_dedup):
    """SymmetryTable.analyze_game_tree (game_stats.py:276-296) with the frontier PARTITIONED over the ranks by
    owner_rank: per depth every rank expands + canonicalises + merges its share (qk_bfs_level), sends each child
    class to its owner (exchange_by_owner), the owner merges (qk_canon_dedup), and one all-reduce sums the
    level's counters.  Same rows as batch.symmetry_table on every rank."""
    import torch
    rank, world = _world(group)
    d = 0 if device is None else int(device)
    on_device = on_device and torch.cuda.is_available()
    empty = np.zeros((1, 8), np.uint16)
    mine = int(owner_rank(empty, world)[0]) == rank
    frontier = empty[: 1 if mine else 0]
    mult = np.ones(1 if mine else 0, np.uint64)
    if on_device:
        frontier = torch.from_numpy(frontier.view(np.int16)).to(f"cuda:{d}")
        mult = torch.from_numpy(mult.astype(np.int64)).to(f"cuda:{d}")
    rows = []
    growth = 64.0                   # local children per local parent at the previous depth: sizes the merge table
    for _depth in range(1, max_depth + 1):
        n = int(frontier.shape[0])
        stats = np.zeros(7, np.int64)
        kids, kmult = np.zeros((0, 8), np.uint16), np.zeros(0, np.uint64)
        if n:
            cap = int(min(64 * n, max(1 << 12, B.level_margin(n) * growth * n)))      # bfs_level doubles it when it was too small
            kids, kmult, st = bfs_fn(frontier, mult, cap=cap, sort_output=False, device=d)
            growth = max(0.05, int(kids.shape[0]) / n)
            stats[:5] = [st["total_moves"], st["line_wins_p0"], st["line_wins_p1"], st["blocked_p0"], st["blocked_p1"]]
        host = not (isinstance(kids, torch.Tensor) and kids.is_cuda)
        if host and on_device:      # an empty shard still has to join the collective on the GPU
            kids = torch.zeros((0, 8), dtype=torch.int16, device=f"cuda:{d}")
            kmult = torch.zeros(0, dtype=torch.int64, device=f"cuda:{d}")
            host = False
        rp, rm = exchange_by_owner(kids, kmult, group)
        if world > 1 and len(rp):
            frontier, mult = dedup_fn(_like_states(rp, host), rm if not host else rm.cpu().numpy().astype(np.uint64),
                                      device=d, sort_output=False)
        elif world > 1:
            frontier, mult = kids[:0], kmult[:0]
        else:
            frontier, mult = kids, kmult
        stats[5] = int(frontier.shape[0])
        stats[6] = int(mult.sum()) if len(mult) else 0
        tot = all_reduce_counts(stats, d, group)
        rows.append([int(tot[0]), int(tot[5]), int(tot[1] + tot[4]), int(tot[2] + tot[3]), int(tot[6])])
    return rows


# ----------------------------------------------------------------------------- the exchange fused into the merge
# Same two jobs as above, but the all-to-all is not a separate collective: the emit pass of the local merge
# stores every distinct class straight into its owner's INBOX over NVLink (include/quantik_b200.h,
# qk_*_scatter / qk_inbox_merge).  torch symmetric memory only provides the peer mapping of the inboxes;
# one small all-reduce per step is both the barrier of the protocol and the "did every region fit" vote.
class PeerInboxes:
    """`copies` inboxes per rank (alternate them between consecutive steps), each mapped into every peer."""

    def __init__(self, slots_per_src: int, device: int, group=None, copies: int = 2):
        self.device, self.group, self.copies = int(device), group, copies
        self.rank, self.world = _world(group)
        self.slots_per_src = 0
        self.bufs, self.ptrs = [], []
        self.ensure(slots_per_src)

    def ensure(self, slots_per_src: int) -> None:
        """Collective: (re)allocate when the regions are smaller than `slots_per_src` (same value on every rank)."""
        import torch
        if slots_per_src <= self.slots_per_src:
            return
        self.bufs, self.ptrs = [], []
        self.slots_per_src = int(slots_per_src)
        nbytes = B.inbox_bytes(self.world, self.slots_per_src)
        for _ in range(self.copies):
            if self.world == 1:
                t = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{self.device}")
                ptrs = [t.data_ptr()]
            else:
                import torch.distributed as dist
                import torch.distributed._symmetric_memory as symm_mem
                t = symm_mem.empty(nbytes, dtype=torch.uint8, device=f"cuda:{self.device}")
                h = symm_mem.rendezvous(t, group=self.group if self.group is not None else dist.group.WORLD)
                ptrs = [int(p) for p in h.buffer_ptrs]
            self.bufs.append(t)
            self.ptrs.append(ptrs)

    def received(self, k: int) -> int:
        """Entries the peers wrote into my inbox k (its header; valid after the barrier)."""
        import torch
        return int(self.bufs[k][: 8 * self.world].view(torch.int64).sum().item())


def _vote(failed: bool, device: int, group=None) -> bool:
    """All-reduce(max) of a flag: the barrier between scatter and merge, and the agreement to retry."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return failed
    t = torch.tensor([int(failed)], dtype=torch.int32, device=f"cuda:{device}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return bool(t.item())


def _max_over_ranks(value: int, device: int, group=None) -> int:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return int(value)
    t = torch.tensor([int(value)], dtype=torch.int64, device=f"cuda:{device}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(t.item())


def fused_canon_dedup(states, mult=None, color_swap: bool = False, device: int = 0, group=None,
                      inboxes: Optional[PeerInboxes] = None, sort_output: bool = False):
    """sharded_canon_dedup with the exchange fused into the merge (GPU only).  Same return value."""
    from ._native import QuantikB200Error
    rank, world = _world(group)
    n = len(states)
    want = _max_over_ranks(int(1.25 * n / world) + 4096, device, group)     # every local state distinct, evenly owned
    if inboxes is None:
        inboxes = PeerInboxes(want, device, group, copies=1)
    inboxes.ensure(want)
    while True:
        failed = False
        try:
            B.canon_dedup_scatter(states, mult, inboxes.ptrs[0], rank, inboxes.slots_per_src, color_swap=color_swap, device=device)
        except QuantikB200Error as e:
            if e.code != -4:
                raise
            failed = True
        if not _vote(failed, device, group):
            break
        inboxes.ensure(2 * inboxes.slots_per_src)
    got = inboxes.received(0)
    mine, mm = B.inbox_merge(inboxes.ptrs[0][rank], world, inboxes.slots_per_src, max(got, 1), device=device, sort_output=sort_output)
    total = all_reduce_counts(np.array([int(mine.shape[0])], np.int64), device, group)
    return mine, mm, int(total[0])


def fused_symmetry_table(max_depth: int, device: int = 0, group=None, inboxes: Optional[PeerInboxes] = None):
    """sharded_symmetry_table with the per-depth exchange fused into the expansion's merge: per depth one
    qk_bfs_level_scatter, one vote (the barrier), one qk_inbox_merge; the counters are all-reduced once at the end."""
    import torch
    from ._native import QuantikB200Error
    rank, world = _world(group)
    d = int(device)
    if inboxes is None:
        inboxes = PeerInboxes(1 << 14, d, group, copies=2)
    # the empty board belongs to whoever owns it under the library's hash: let the merge decide -- rank 0 scatters it
    frontier = torch.zeros((1 if rank == 0 else 0, 8), dtype=torch.int16, device=f"cuda:{d}")
    mult = torch.ones(1 if rank == 0 else 0, dtype=torch.int64, device=f"cuda:{d}")
    B.canon_dedup_scatter(frontier, mult, inboxes.ptrs[1], rank, inboxes.slots_per_src, device=d)
    _vote(False, d, group)
    frontier, mult = B.inbox_merge(inboxes.ptrs[1][rank], world, inboxes.slots_per_src, 1, device=d)
    levels = np.zeros((max_depth, 7), np.int64)
    growth = 64.0
    for depth in range(1, max_depth + 1):
        k = depth % 2
        n = int(frontier.shape[0])
        local_cap = int(min(64 * max(n, 1), max(1 << 12, B.level_margin(n) * growth * n)))
        want = _max_over_ranks(int(1.25 * local_cap / B.level_margin(n) / world) + 4096, d, group)
        inboxes.ensure(want)
        while True:
            failed = False
            try:
                st, sent = B.bfs_level_scatter(frontier, mult, inboxes.ptrs[k], rank, inboxes.slots_per_src, local_cap, device=d)
            except QuantikB200Error as e:
                if e.code != -4:
                    raise
                failed = True
            if not _vote(failed, d, group):
                break
            local_cap *= 2                      # whichever of the two was short, on whichever rank
            inboxes.ensure(2 * inboxes.slots_per_src)
        growth = max(0.05, float(sent.sum()) / max(n, 1))
        got = inboxes.received(k)
        frontier, mult = B.inbox_merge(inboxes.ptrs[k][rank], world, inboxes.slots_per_src, max(got, 1), device=d)
        levels[depth - 1, :5] = [st["total_moves"], st["line_wins_p0"], st["line_wins_p1"], st["blocked_p0"], st["blocked_p1"]]
        levels[depth - 1, 5] = int(frontier.shape[0])
        levels[depth - 1, 6] = int(mult.sum()) if len(mult) else 0
    tot = all_reduce_counts(levels, d, group)
    return [[int(t[0]), int(t[5]), int(t[1] + t[4]), int(t[2] + t[3]), int(t[6])] for t in tot]
