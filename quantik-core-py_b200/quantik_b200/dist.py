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


def partition_by_cost(cost: np.ndarray, world: int):
    """Static partition of weighted roots ("LPT/greedy by probe cost"): roots in largest-first order of
    their probe cost (ties by index), dealt in serpentine order 0..w-1, w-1..0, ... -- the vectorised form of
    greedy longest-processing-time assignment: after every two rounds all ranks hold nearly equal sums.
    -> list of `world` index arrays (sorted ascending).  Same answer on every rank (pure function of cost)."""
    cost = np.asarray(cost)
    order = np.argsort(-cost.astype(np.int64), kind="stable")
    pos = np.arange(len(order)) % (2 * world)
    dest = np.where(pos < world, pos, 2 * world - 1 - pos)
    return [np.sort(order[dest == r]) for r in range(world)]


def sharded_perft(roots: np.ndarray, depth: int, mult: Optional[np.ndarray] = None, device: Optional[int] = None,
                  group=None, perft_fn: Callable[..., Dict[str, np.ndarray]] = B.perft,
                  probe_fn: Callable[..., np.ndarray] = B.perft_probe) -> Dict[str, np.ndarray]:
    """BASELINE configs 2/5: the weighted roots are dealt to the ranks largest-first by a depth-2 probe of their
    subtrees (partition_by_cost; the reference chunks its frontier by count, examples/generate_opening_book.py:294-332),
    every rank enumerates below its roots -- inside the GPU the depth-first kernels pull (root, child) items from one
    atomic queue -- and ONE all-reduce of the 5 x (depth+1) counters closes the job."""
    rank, world = _world(group)
    roots = np.asarray(roots, np.uint16).reshape(-1, 8)
    if world > 1 and len(roots):
        idx = partition_by_cost(probe_fn(roots, device=device), world)[rank]
    else:
        idx = np.arange(len(roots))
    mine = np.ascontiguousarray(roots[idx])
    m = None if mult is None else np.ascontiguousarray(np.asarray(mult, np.uint64)[idx])
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


def _mix64(h, l):
    """qk_dedup.cu `mix` on torch int64 tensors (two's complement wrap-around = the kernel's u64 arithmetic)."""
    x = (h * -7046029254386353131) ^ ((l + -2960836687051489901) * -4417276706812531889)
    x = x ^ ((x >> 29) & 0x7FFFFFFFF)                 # logical shift
    x = x * -4658895280553007687
    return x ^ ((x >> 32) & 0xFFFFFFFF)


def owner_rank(payloads, world: int):
    """Owner of each canonical payload -- the SAME function as the library's owner_of (csrc/qk_dedup.cu): the two
    big-endian 64-bit halves of the 16 payload bytes, mixed, high 32 bits scaled to [0, world).  Both exchange
    implementations (NCCL all-to-all here, peer-memory scatter in the library) therefore partition the classes
    over the ranks identically (torch int64 tensor)."""
    import torch
    p = _as_i64_pairs(payloads)
    be = p.view(torch.uint8).reshape(-1, 2, 8).flip(-1).contiguous().view(torch.int64).reshape(-1, 2)
    top = (_mix64(be[:, 1], be[:, 0]) >> 32) & 0xFFFFFFFF
    return (top * int(world)) >> 32


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


VOTE_OK, VOTE_RETRY, VOTE_FATAL = 0, 1, 2


def _vote(status: int, device: int, group=None, extra: Tuple[int, ...] = ()) -> Tuple[int, ...]:
    """ONE all-reduce(max) that is (a) the barrier between scatter and merge, (b) the agreement on what happens
    next -- VOTE_OK, VOTE_RETRY (some region or table was too small) or VOTE_FATAL (some rank failed: every rank
    raises after the collective instead of one raising before it and the others waiting forever) -- and
    (c) carries `extra` non-negative integers, each reduced with max.  -> (status, *extra) over all ranks."""
    import torch
    import torch.distributed as dist
    vals = [int(status)] + [int(x) for x in extra]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return tuple(vals)
    t = torch.tensor(vals, dtype=torch.int64, device=f"cuda:{device}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return tuple(int(v) for v in t.tolist())


def _guarded(fn):
    """Run one local step of a fused protocol.  -> (result, vote status, exception): a short region / table is a
    retry (QK_E_NOMEM), anything else is fatal -- but the caller still enters the vote."""
    from ._native import QuantikB200Error
    try:
        return fn(), VOTE_OK, None
    except QuantikB200Error as e:
        return None, (VOTE_RETRY if e.code == -4 else VOTE_FATAL), e
    except Exception as e:                      # noqa: BLE001 -- reported to every rank through the vote
        return None, VOTE_FATAL, e


def _raise_if_fatal(status: int, err, what: str) -> None:
    if status >= VOTE_FATAL:
        if err is not None:
            raise err
        raise RuntimeError(f"{what}: another rank failed (see its log); aborting on every rank")


def fused_canon_dedup(states, mult=None, color_swap: bool = False, device: int = 0, group=None,
                      inboxes: Optional[PeerInboxes] = None, sort_output: bool = False, n_max: Optional[int] = None):
    """sharded_canon_dedup with the exchange fused into the merge (GPU only): same partition of the classes over
    the ranks (owner_rank = the library's owner_of), so the same (payloads this rank owns, their global
    multiplicities, global number of distinct classes) up to the order of the rows.
    `n_max` = the largest len(states) over the ranks when the caller knows it (saves the sizing all-reduce)."""
    rank, world = _world(group)
    n = len(states)
    if n_max is None:
        n_max = _vote(VOTE_OK, device, group, (n,))[1]
    want = int(1.25 * n_max / world) + 4096                     # every local state distinct, evenly owned
    if inboxes is None:
        inboxes = PeerInboxes(want, device, group, copies=1)
    inboxes.ensure(want)
    while True:
        _, status, err = _guarded(lambda: B.canon_dedup_scatter(states, mult, inboxes.ptrs[0], rank, inboxes.slots_per_src,
                                                                 color_swap=color_swap, device=device))
        status = _vote(status, device, group)[0]
        _raise_if_fatal(status, err, "fused_canon_dedup")
        if status == VOTE_OK:
            break
        inboxes.ensure(2 * inboxes.slots_per_src)
    got = inboxes.received(0)
    res, status, err = _guarded(lambda: B.inbox_merge(inboxes.ptrs[0][rank], world, inboxes.slots_per_src, max(got, 1),
                                                      device=device, sort_output=sort_output))
    n_mine = int(res[0].shape[0]) if res is not None else 0
    total = all_reduce_counts(np.array([n_mine, int(status != VOTE_OK)], np.int64), device, group)
    _raise_if_fatal(VOTE_FATAL if total[1] else VOTE_OK, err, "fused_canon_dedup (merge)")
    return res[0], res[1], int(total[0])


def fused_symmetry_table(max_depth: int, device: int = 0, group=None, inboxes: Optional[PeerInboxes] = None):
    """sharded_symmetry_table with the per-depth exchange fused into the expansion's merge: per depth one
    qk_bfs_level_scatter, ONE collective (the vote: barrier + retry agreement + the sizes the next depth needs)
    and one qk_inbox_merge; the counters are all-reduced once at the end.  Inboxes alternate between consecutive
    depths and nothing else synchronises the ranks, so a fast rank delivers depth d+1 while a slow one still
    merges depth d."""
    import torch
    rank, world = _world(group)
    d = int(device)
    if inboxes is None:
        inboxes = PeerInboxes(1 << 14, d, group, copies=2)
    # the empty board belongs to whoever owns it under the library's hash: let the merge decide -- rank 0 scatters it
    frontier = torch.zeros((1 if rank == 0 else 0, 8), dtype=torch.int16, device=f"cuda:{d}")
    mult = torch.ones(1 if rank == 0 else 0, dtype=torch.int64, device=f"cuda:{d}")
    _, status, err = _guarded(lambda: B.canon_dedup_scatter(frontier, mult, inboxes.ptrs[1], rank, inboxes.slots_per_src, device=d))
    _raise_if_fatal(_vote(status, d, group)[0], err, "fused_symmetry_table (root)")
    frontier, mult = B.inbox_merge(inboxes.ptrs[1][rank], world, inboxes.slots_per_src, 1, device=d)
    levels = np.zeros((max_depth, 7), np.int64)
    growth = 64.0                   # children sent per parent at the previous depth, max over ranks (x 1024 on the wire)
    n_max = 1                       # largest frontier share at this depth, max over ranks
    carry, carry_err = VOTE_OK, None            # a failed merge is reported through the NEXT vote
    for depth in range(1, max_depth + 1):
        k = depth % 2
        n = int(frontier.shape[0])
        local_cap = int(min(64 * max(n, 1), max(1 << 12, B.level_margin(n) * growth * max(n, 1))))
        # regions: what the busiest rank will send, spread evenly over the owners (a short region is a retry)
        inboxes.ensure(int(1.25 * min(64.0, growth) * n_max / world) + 4096)
        while True:
            if carry == VOTE_OK:
                res, status, err = _guarded(lambda: B.bfs_level_scatter(frontier, mult, inboxes.ptrs[k], rank,
                                                                        inboxes.slots_per_src, local_cap, device=d))
            else:
                res, status, err = None, carry, carry_err
            sent_total = int(res[1].sum()) if res is not None else 0
            status, sent_max, ratio_max = _vote(status, d, group, (sent_total, int(1024.0 * sent_total / max(n, 1)) + 1))
            _raise_if_fatal(status, err, f"fused_symmetry_table (depth {depth})")
            if status == VOTE_OK:
                break
            local_cap *= 2                      # whichever of the two was short, on whichever rank
            inboxes.ensure(2 * inboxes.slots_per_src)
        st = res[0]
        growth = max(0.05, ratio_max / 1024.0)
        n_max = max(1, int(1.1 * sent_max))     # owners are hash-uniform: no rank receives much more than the busiest sent
        merged, carry, carry_err = _guarded(lambda: B.inbox_merge(inboxes.ptrs[k][rank], world, inboxes.slots_per_src,
                                                                  max(inboxes.received(k), 1), device=d))
        if merged is None:
            carry = VOTE_FATAL                  # even a "retry" code is fatal here: the peers' data is already delivered
            continue
        frontier, mult = merged
        levels[depth - 1, :5] = [st["total_moves"], st["line_wins_p0"], st["line_wins_p1"], st["blocked_p0"], st["blocked_p1"]]
        levels[depth - 1, 5] = int(frontier.shape[0])
        levels[depth - 1, 6] = int(mult.sum()) if len(mult) else 0
    status = _vote(carry, d, group)[0] if carry != VOTE_OK or world > 1 else VOTE_OK
    _raise_if_fatal(status, carry_err, "fused_symmetry_table (last merge)")
    tot = all_reduce_counts(levels, d, group)
    return [[int(t[0]), int(t[5]), int(t[1] + t[4]), int(t[2] + t[3]), int(t[6])] for t in tot]
