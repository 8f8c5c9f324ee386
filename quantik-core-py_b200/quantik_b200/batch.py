"""Batched hot-path operations over N x 8 uint16 state buffers.

Inputs may be numpy arrays (host: staged by the library, call returns when the
result is in place) or CUDA torch tensors (device: zero-copy, asynchronous on
torch's current stream).  Outputs come back in the same kind of container.

The buffer layout is the reference's batch wire format
(``storage/compact_state.py:125-147``: N concatenated ``<8H`` payloads).
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Tuple

import numpy as np

from . import _native as N

SEED_DEFAULT = 0x5155414E54494B31  # "QUANTIK1"

try:  # torch is plumbing only (device buffers, streams); numpy-only use works without it
    import torch
except Exception:  # pragma: no cover
    torch = None


def _is_torch(x) -> bool:
    return torch is not None and isinstance(x, torch.Tensor)


class _Buf:
    """pointer + owner of one argument"""

    __slots__ = ("ptr", "obj", "n")

    def __init__(self, ptr, obj, n):
        self.ptr, self.obj, self.n = ptr, obj, n


def _states_in(x) -> Tuple[_Buf, bool, int]:
    """-> (buffer, on_device, device ordinal)"""
    if _is_torch(x):
        if not x.is_cuda:
            return _states_in(x.numpy())
        if not x.is_contiguous():
            x = x.contiguous()
        nbytes = x.numel() * x.element_size()
        if nbytes % 16:
            raise ValueError("state tensor must hold a whole number of 16-byte states")
        return _Buf(x.data_ptr(), x, nbytes // 16), True, x.device.index or 0
    a = np.ascontiguousarray(x, dtype=np.uint16)
    if a.ndim == 1:
        if a.size % 8:
            raise ValueError("Invalid bitboard data")  # core.py:30-31
        a = a.reshape(-1, 8)
    if a.ndim != 2 or a.shape[1] != 8:
        raise ValueError("states must have shape (N, 8)")
    return _Buf(a.ctypes.data, a, a.shape[0]), False, 0


def _out(shape, dtype, like_dev: bool, dev: int):
    if like_dev:
        tdt = {np.uint8: torch.uint8, np.int8: torch.int8, np.uint16: torch.int16, np.uint32: torch.int32,
               np.uint64: torch.int64}[dtype]
        t = torch.empty(shape, dtype=tdt, device=f"cuda:{dev}")
        return _Buf(t.data_ptr(), t, t.shape[0] if t.ndim else 1)
    a = np.empty(shape, dtype)
    return _Buf(a.ctypes.data, a, a.shape[0] if a.ndim else 1)


def _aux_in(x, dtype, on_dev: bool, dev: int) -> _Buf:
    if x is None:
        return _Buf(None, None, 0)
    if _is_torch(x) and x.is_cuda:
        x = x.contiguous()
        return _Buf(x.data_ptr(), x, x.numel())
    a = np.ascontiguousarray(x.cpu().numpy() if _is_torch(x) else x, dtype=dtype)
    return _Buf(a.ctypes.data, a, a.size)


def _stream(on_dev: bool, dev: int):
    if on_dev:
        return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    return None


def _dev(device: Optional[int], on_dev: bool, dev: int) -> int:
    return dev if on_dev else (0 if device is None else int(device))


# ----------------------------------------------------------------------------- primitives
def movegen_masks(states, device: Optional[int] = None):
    """generate_legal_moves for every state -> (action_mask u64, to_move u8, status u8).

    Replaces ``move.generate_legal_moves`` (move.py:199-268): bit ``shape*16+pos`` of the mask
    is a legal move of ``to_move``; an invalid state gives mask 0 / to_move 0 (move.py:223-225)
    with ``status`` = the ``ValidationResult`` code (state_validator.py:16-27)."""
    s, on_dev, d = _states_in(states)
    d = _dev(device, on_dev, d)
    lib = N.init(d)
    mask, tm, st = _out((s.n,), np.uint64, on_dev, d), _out((s.n,), np.uint8, on_dev, d), _out((s.n,), np.uint8, on_dev, d)
    N.check(lib.qk_movegen_masks(s.ptr, s.n, mask.ptr, tm.ptr, st.ptr, d, _stream(on_dev, d)))
    return mask.obj, tm.obj, st.obj


def apply_moves(states, actions, device: Optional[int] = None):
    """apply_move (move.py:135-166) of action ``shape*16+pos`` by the side to move."""
    s, on_dev, d = _states_in(states)
    d = _dev(device, on_dev, d)
    lib = N.init(d)
    a = _aux_in(actions, np.uint8, on_dev, d)
    if a.n != s.n:
        raise ValueError("one action per state required")
    out = _out((s.n, 8), np.uint16, on_dev, d)
    N.check(lib.qk_apply_moves(s.ptr, a.ptr, s.n, out.ptr, d, _stream(on_dev, d)))
    return out.obj


def win_status(states, device: Optional[int] = None):
    """check_game_winner (game_utils.py:161-189) -> WinStatus codes."""
    s, on_dev, d = _states_in(states)
    d = _dev(device, on_dev, d)
    lib = N.init(d)
    out = _out((s.n,), np.uint8, on_dev, d)
    N.check(lib.qk_win_status(s.ptr, s.n, out.ptr, d, _stream(on_dev, d)))
    return out.obj


def canonical_payloads(states, color_swap: bool = False, want_transform: bool = False, device: Optional[int] = None):
    """find_canonical_form (symmetry.py:289-354) -> (N, 8) canonical bitboards
    [, transform codes d4 | cs<<3 | perm[i] << (4+2i)]."""
    s, on_dev, d = _states_in(states)
    d = _dev(device, on_dev, d)
    lib = N.init(d)
    out = _out((s.n, 8), np.uint16, on_dev, d)
    tr = _out((s.n,), np.uint16, on_dev, d) if want_transform else _Buf(None, None, 0)
    N.check(lib.qk_canonical(s.ptr, s.n, int(bool(color_swap)), out.ptr, tr.ptr, d, _stream(on_dev, d)))
    return (out.obj, tr.obj) if want_transform else out.obj


def canonical_keys(states, device: Optional[int] = None):
    """State.canonical_key (core.py:171-180) for a host batch -> list of 18-byte keys."""
    pay = canonical_payloads(np.asarray(states, dtype=np.uint16), device=device)
    raw = np.ascontiguousarray(pay).astype("<u2").tobytes()
    return [b"\x01\x02" + raw[16 * i:16 * i + 16] for i in range(len(pay))]


def orbit_sizes(states, device: Optional[int] = None):
    """count_orbit_size (symmetry.py:356-396)."""
    s, on_dev, d = _states_in(states)
    d = _dev(device, on_dev, d)
    lib = N.init(d)
    out = _out((s.n,), np.uint8, on_dev, d)
    N.check(lib.qk_orbit_size(s.ptr, s.n, out.ptr, d, _stream(on_dev, d)))
    return out.obj


def encode_transform(d4: int, color_swap: bool, perm) -> int:
    return int(d4) | (int(bool(color_swap)) << 3) | sum(int(p) << (4 + 2 * i) for i, p in enumerate(perm))


def decode_transform(code: int):
    code = int(code)
    return code & 7, bool(code & 8), tuple((code >> (4 + 2 * i)) & 3 for i in range(4))


def apply_symmetry(states, transforms, device: Optional[int] = None):
    """apply_symmetry (symmetry.py:252-287) with one transform code per state."""
    s, on_dev, d = _states_in(states)
    d = _dev(device, on_dev, d)
    lib = N.init(d)
    t = _aux_in(transforms, np.uint16, on_dev, d)
    if t.n != s.n:
        raise ValueError("one transform per state required")
    out = _out((s.n, 8), np.uint16, on_dev, d)
    N.check(lib.qk_apply_symmetry(s.ptr, t.ptr, s.n, out.ptr, d, _stream(on_dev, d)))
    return out.obj


def apply_symmetry_to_moves(players, actions, transforms, inverse: bool = False, device: Optional[int] = None):
    """apply_symmetry_to_move (symmetry.py:425-464) for a batch: move i = (players[i], actions[i] = shape * 16 + square)
    under transform code i (encode_transform).  inverse=True applies SymmetryTransform.inverse() of every code: with the
    codes `canonical_payloads(..., want_transform=True)` returns, that takes moves found on canonical positions back to
    the positions that were looked up.  -> (players, actions) uint8."""
    on_dev = _is_torch(actions) and actions.is_cuda
    d = _dev(device, on_dev, actions.device.index if on_dev else 0)
    lib = N.init(d)
    a = _aux_in(actions, np.uint8, on_dev, d)
    pl = _aux_in(players, np.uint8, on_dev, d)
    t = _aux_in(transforms, np.uint16, on_dev, d)
    if not (a.n == pl.n == t.n):
        raise ValueError("one player and one transform per action required")
    op, oa = _out((a.n,), np.uint8, on_dev, d), _out((a.n,), np.uint8, on_dev, d)
    N.check(lib.qk_apply_symmetry_to_moves(pl.ptr, a.ptr, t.ptr, a.n, int(bool(inverse)), op.ptr, oa.ptr, d, _stream(on_dev, d)))
    return op.obj, oa.obj


def canon_dedup(states, mult=None, color_swap: bool = False, cap: Optional[int] = None, device: Optional[int] = None,
                sort_output: bool = True):
    """Canonicalise, merge equal payloads, sum multiplicities (game_stats.py:487-505).
    -> (unique payloads (U, 8) -- sorted in byte order unless sort_output=False --, multiplicities (U,))."""
    s, on_dev, d = _states_in(states)
    d = _dev(device, on_dev, d)
    lib = N.init(d)
    m = _aux_in(mult, np.uint64, on_dev, d)
    cap = s.n if cap is None else int(cap)
    uniq = _out((cap + 1, 8), np.uint16, on_dev, d)
    um = _out((cap + 1,), np.uint64, on_dev, d)
    nu = ctypes.c_size_t(0)
    flags = int(bool(color_swap)) | (0 if sort_output else 2)          # QK_DEDUP_UNSORTED
    N.check(lib.qk_canon_dedup(s.ptr, m.ptr, s.n, flags, uniq.ptr, um.ptr, cap, ctypes.byref(nu),
                               d, _stream(on_dev, d)))
    return uniq.obj[:nu.value], um.obj[:nu.value]


def level_margin(n_parents: int) -> float:
    """Head-room of the next frontier's size estimate (children per parent of the previous depth x parents): the
    ratio only falls as the board fills up, so big late levels need little; a level that still overflows is
    repeated with a doubled table (QK_E_NOMEM)."""
    return 1.6 if n_parents < (1 << 20) else 1.25


def bfs_level(states, mult=None, color_swap: bool = False, cap: int = 1 << 20, sort_output: bool = True,
              device: Optional[int] = None):
    """One depth of the reference's SymmetryTable / opening-book BFS (game_stats.py:359-485), fused on the
    GPU: expand every parent, drop line-completing moves, canonicalise + merge the rest.
    -> (next frontier (U, 8) [sorted in byte order], multiplicities (U,), stats dict)."""
    s, on_dev, d = _states_in(states)
    d = _dev(device, on_dev, d)
    lib = N.init(d)
    m = _aux_in(mult, np.uint64, on_dev, d)
    while True:
        uniq = _out((cap + 1, 8), np.uint16, on_dev, d)
        um = _out((cap + 1,), np.uint64, on_dev, d)
        nu = ctypes.c_size_t(0)
        stats = np.zeros(5, np.uint64)
        rc = lib.qk_bfs_level(s.ptr, m.ptr, s.n, int(bool(color_swap)), int(bool(sort_output)), uniq.ptr, um.ptr, cap,
                              ctypes.byref(nu), stats.ctypes.data, d, _stream(on_dev, d))
        if rc == -4:            # QK_E_NOMEM: frontier larger than cap -> retry with a bigger table
            del uniq, um
            cap *= 2
            continue
        N.check(rc)
        break
    names = ["total_moves", "line_wins_p0", "line_wins_p1", "blocked_p0", "blocked_p1"]
    return uniq.obj[:nu.value], um.obj[:nu.value], {k: int(v) for k, v in zip(names, stats)}


def symmetry_table(max_depth: int, device: Optional[int] = None, on_device: bool = True, sort_output: bool = False,
                   verbose: bool = False):
    """The reference's SymmetryTable.analyze_game_tree (game_stats.py:276-296) from the empty board:
    rows [total_legal_moves, unique_canonical, p0_wins, p1_wins, ongoing] for depth 1..max_depth, in the
    reference's convention (a blocked mover is booked as a win of the other player, :397-398)."""
    import time as _time
    d = 0 if device is None else int(device)
    frontier, mult = np.zeros((1, 8), np.uint16), np.ones(1, np.uint64)
    if on_device and torch is not None:
        frontier = torch.zeros((1, 8), dtype=torch.int16, device=f"cuda:{d}")
        mult = torch.ones(1, dtype=torch.int64, device=f"cuda:{d}")
    rows = []
    growth = 64.0
    for depth in range(1, max_depth + 1):
        n = int(frontier.shape[0])
        if n == 0:
            rows.append([0, 0, 0, 0, 0])
            continue
        t0 = _time.perf_counter()
        cap = int(min(64 * n, max(1 << 12, level_margin(n) * growth * n)))
        nxt, nmult, st = bfs_level(frontier, mult, cap=cap, sort_output=sort_output, device=d)
        del frontier, mult
        frontier, mult = nxt, nmult
        ongoing = int(mult.sum()) if len(mult) else 0
        rows.append([st["total_moves"], int(frontier.shape[0]), st["line_wins_p0"] + st["blocked_p1"],
                     st["line_wins_p1"] + st["blocked_p0"], ongoing])
        growth = max(0.05, int(frontier.shape[0]) / n)
        if verbose:
            print(f"depth {depth:2d}: {rows[-1]}  ({_time.perf_counter() - t0:.2f} s)", flush=True)
    return rows


# ----------------------------------------------------------------------------- exchange step over peer memory
def inbox_bytes(world: int, slots_per_src: int) -> int:
    """Size of one rank's inbox (qk_inbox_bytes)."""
    b = ctypes.c_size_t(0)
    N.check(N.load().qk_inbox_bytes(int(world), int(slots_per_src), ctypes.byref(b)))
    return int(b.value)


def payload_owner(payloads, world: int, device: Optional[int] = None):
    """Owner rank of each canonical payload (qk_payload_owner) -> uint8 per payload."""
    s, on_dev, d = _states_in(payloads)
    d = _dev(device, on_dev, d)
    lib = N.init(d)
    out = _out((s.n,), np.uint8, on_dev, d)
    N.check(lib.qk_payload_owner(s.ptr, s.n, int(world), out.ptr, d, _stream(on_dev, d)))
    return out.obj


def _ptr_array(ptrs):
    return (ctypes.c_void_p * len(ptrs))(*[int(p) for p in ptrs])


def canon_dedup_scatter(states, mult, inbox_ptrs, rank: int, slots_per_src: int, color_swap: bool = False,
                        device: Optional[int] = None) -> np.ndarray:
    """Canonicalise + merge locally, then store every distinct class into its owner's inbox over peer memory
    (qk_canon_dedup_scatter).  -> entries written per destination rank."""
    s, on_dev, d = _states_in(states)
    d = _dev(device, on_dev, d)
    lib = N.init(d)
    m = _aux_in(mult, np.uint64, on_dev, d)
    sent = np.zeros(len(inbox_ptrs), np.uint64)
    N.check(lib.qk_canon_dedup_scatter(s.ptr, m.ptr, s.n, int(bool(color_swap)), _ptr_array(inbox_ptrs), len(inbox_ptrs),
                                       int(rank), int(slots_per_src), sent.ctypes.data, d, _stream(on_dev, d)))
    return sent


def bfs_level_scatter(states, mult, inbox_ptrs, rank: int, slots_per_src: int, local_cap: int, color_swap: bool = False,
                      device: Optional[int] = None):
    """One BFS depth of this rank's share of the frontier, children delivered to their owners' inboxes
    (qk_bfs_level_scatter).  -> (stats dict, entries written per destination)."""
    s, on_dev, d = _states_in(states)
    d = _dev(device, on_dev, d)
    lib = N.init(d)
    m = _aux_in(mult, np.uint64, on_dev, d)
    sent = np.zeros(len(inbox_ptrs), np.uint64)
    stats = np.zeros(5, np.uint64)
    N.check(lib.qk_bfs_level_scatter(s.ptr, m.ptr, s.n, int(bool(color_swap)), int(local_cap), _ptr_array(inbox_ptrs),
                                     len(inbox_ptrs), int(rank), int(slots_per_src), stats.ctypes.data, sent.ctypes.data,
                                     d, _stream(on_dev, d)))
    names = ["total_moves", "line_wins_p0", "line_wins_p1", "blocked_p0", "blocked_p1"]
    return {k: int(v) for k, v in zip(names, stats)}, sent


def inbox_merge(inbox_ptr: int, world: int, slots_per_src: int, cap: int, device: int = 0, sort_output: bool = False):
    """Merge what the peers delivered into this rank's inbox (qk_inbox_merge; call after a barrier).
    -> (payloads (U, 8) int16 CUDA tensor, multiplicities (U,) int64 CUDA tensor)."""
    lib = N.init(device)
    uniq = _out((cap + 1, 8), np.uint16, True, device)
    um = _out((cap + 1,), np.uint64, True, device)
    nu = ctypes.c_size_t(0)
    N.check(lib.qk_inbox_merge(ctypes.c_void_p(int(inbox_ptr)), int(world), int(slots_per_src), int(bool(sort_output)),
                               uniq.ptr, um.ptr, int(cap), ctypes.byref(nu), device, _stream(True, device)))
    return uniq.obj[:nu.value], um.obj[:nu.value]


def book_level(states, color_swap: bool = False, cap: int = 1 << 20, device: Optional[int] = None):
    """One depth of the opening-book BFS (examples/generate_opening_book.py:185-280): finished classes are
    terminal and not expanded, every other class contributes its distinct canonical children.
    -> (classes of the next depth (U, 8), their in-degree (U,), {"terminal": ..., "edges": ...})."""
    s, on_dev, d = _states_in(states)
    d = _dev(device, on_dev, d)
    lib = N.init(d)
    while True:
        uniq = _out((cap + 1, 8), np.uint16, on_dev, d)
        deg = _out((cap + 1,), np.uint64, on_dev, d)
        nu = ctypes.c_size_t(0)
        stats = np.zeros(2, np.uint64)
        rc = lib.qk_book_level(s.ptr, s.n, int(bool(color_swap)), uniq.ptr, deg.ptr, cap, ctypes.byref(nu),
                               stats.ctypes.data, d, _stream(on_dev, d))
        if rc == -4:
            del uniq, deg
            cap *= 2
            continue
        N.check(rc)
        break
    return uniq.obj[:nu.value], deg.obj[:nu.value], {"terminal": int(stats[0]), "edges": int(stats[1])}


def opening_book_summary(depth: int, device: Optional[int] = None) -> dict:
    """The ``opening-book-summary.v1`` record (opening_book_summary.py:36-93; the cross-stack contract
    of .github/workflows/contracts.yml:20-75) of a full-width book built from the empty board to
    ``depth``, computed by the GPU BFS instead of sqlite."""
    if depth < 0:
        raise ValueError("depth must be non-negative")          # opening_book_summary.py:38-39
    d = 0 if device is None else int(device)
    frontier = np.zeros((1, 8), np.uint16)
    if torch is not None:
        frontier = torch.zeros((1, 8), dtype=torch.int16, device=f"cuda:{d}")
    per_depth = []
    growth = 64.0
    for cur in range(depth + 1):
        n = int(frontier.shape[0])
        cap = int(min(64 * max(n, 1), max(1 << 12, level_margin(n) * growth * n)))
        nxt, _deg, st = book_level(frontier, cap=cap, device=d) if n else (frontier, None, {"terminal": 0, "edges": 0})
        per_depth.append({"depth": cur, "positions": n, "terminal": st["terminal"], "edges": st["edges"]})
        growth = max(0.05, int(nxt.shape[0]) / max(n, 1))
        frontier = nxt
    return {"schema": "opening-book-summary.v1", "contract_version": "1.1.0", "depth": depth,
            "total_positions": sum(r["positions"] for r in per_depth),
            "terminal_positions": sum(r["terminal"] for r in per_depth),
            "total_edges": sum(r["edges"] for r in per_depth), "per_depth": per_depth}


# ----------------------------------------------------------------------------- playouts
def playouts(roots, rollouts: int, seed: int = SEED_DEFAULT, max_plies: int = -1, first_rollout: int = 0,
             root_index_base: int = 0, per_playout: bool = False, device: Optional[int] = None) -> Dict[str, object]:
    """``rollouts`` uniform-random playouts from every root.

    max_plies < 0: BeamSearchEngine._rollout semantics (beam_search.py:624-645), results +-1.
    max_plies >= 0: MCTSEngine._simulate semantics (mcts.py:385-437), 0 after max_plies moves.
    Returns wins_p0 / wins_p1 / draws (u32 per root) and, if ``per_playout``, ``outcome``
    (int8, P0 view) and ``plies`` (u8) of shape (n_roots, rollouts)."""
    s, on_dev, d = _states_in(roots)
    d = _dev(device, on_dev, d)
    lib = N.init(d)
    n = s.n
    w0, w1, dr = (_out((n,), np.uint32, on_dev, d) for _ in range(3))
    oc = _out((n, rollouts), np.int8, on_dev, d) if per_playout else _Buf(None, None, 0)
    pl = _out((n, rollouts), np.uint8, on_dev, d) if per_playout else _Buf(None, None, 0)
    N.check(lib.qk_playouts_ex(s.ptr, n, int(rollouts), int(seed) & 0xFFFFFFFFFFFFFFFF, int(first_rollout),
                               int(root_index_base), int(max_plies), w0.ptr, w1.ptr, dr.ptr, oc.ptr, pl.ptr,
                               d, _stream(on_dev, d)))
    out = {"wins_p0": w0.obj, "wins_p1": w1.obj, "draws": dr.obj}
    if per_playout:
        out["outcome"], out["plies"] = oc.obj, pl.obj
    return out


SEEDED_WEIGHTS = (100.0, -100.0, 20.0, 3.0, 2.0, 0.0)     # evaluation.py:40 (_SEEDED_WEIGHTS)
FEATURE_NAMES = ["threat_own", "threat_opp", "threat_shared", "mobility_diff", "build_two", "build_one"]   # evaluation.py:32-39


def eval_features(states, players, device: Optional[int] = None):
    """evaluation.features (evaluation.py:130-194) for a batch of valid positions -> (N, 6) float64."""
    s, on_dev, d = _states_in(states)
    d = _dev(device, on_dev, d)
    lib = N.init(d)
    pl = _aux_in(players, np.uint8, on_dev, d)
    if pl.n != s.n:
        raise ValueError("one player per state required")
    if on_dev:
        t = torch.empty((s.n, 6), dtype=torch.float64, device=f"cuda:{d}")
        out = _Buf(t.data_ptr(), t, s.n)
    else:
        a = np.empty((s.n, 6), np.float64)
        out = _Buf(a.ctypes.data, a, s.n)
    N.check(lib.qk_eval_features(s.ptr, pl.ptr, s.n, out.ptr, d, _stream(on_dev, d)))
    return out.obj


def evaluate(states, players, weights=SEEDED_WEIGHTS, device: Optional[int] = None):
    """evaluation.evaluate (evaluation.py:197-212): weights @ features, one score per position."""
    f = eval_features(states, players, device=device)
    w = np.asarray(weights, np.float64)
    if _is_torch(f):
        return f @ torch.from_numpy(w).to(f.device)
    return f @ w


def guided_playouts(roots, rollouts: int, seed: int = SEED_DEFAULT, epsilon: float = 0.2, weights=SEEDED_WEIGHTS,
                    max_plies: int = -1, first_rollout: int = 0, root_index_base: int = 0, per_playout: bool = False,
                    device: Optional[int] = None) -> Dict[str, object]:
    """Playouts under the reference's eval-guided epsilon-greedy rollout policy
    (MCTSConfig.rollout_eval_config / rollout_epsilon; mcts.py:345-383).  Same outputs as :func:`playouts`."""
    if not 0.0 <= epsilon <= 1.0:
        raise ValueError(f"rollout_epsilon must be in [0.0, 1.0], got {epsilon}")     # mcts.py:70-76
    s, on_dev, d = _states_in(roots)
    d = _dev(device, on_dev, d)
    lib = N.init(d)
    n = s.n
    w = np.ascontiguousarray(weights, dtype=np.float64)
    if w.shape != (6,):
        raise ValueError("weights must have 6 entries (FEATURE_NAMES order)")
    w0, w1, dr = (_out((n,), np.uint32, on_dev, d) for _ in range(3))
    oc = _out((n, rollouts), np.int8, on_dev, d) if per_playout else _Buf(None, None, 0)
    pl = _out((n, rollouts), np.uint8, on_dev, d) if per_playout else _Buf(None, None, 0)
    N.check(lib.qk_playouts_guided(s.ptr, n, int(rollouts), int(seed) & 0xFFFFFFFFFFFFFFFF, int(first_rollout),
                                   int(root_index_base), int(max_plies), float(epsilon), w.ctypes.data, w0.ptr, w1.ptr,
                                   dr.ptr, oc.ptr, pl.ptr, d, _stream(on_dev, d)))
    out = {"wins_p0": w0.obj, "wins_p1": w1.obj, "draws": dr.obj}
    if per_playout:
        out["outcome"], out["plies"] = oc.obj, pl.obj
    return out


def solve(states, device: Optional[int] = None):
    """Exact game value with mate distance (MinimaxEngine.solve, minimax.py:182-204) for late-game positions.
    -> (outcome int8 +1/-1 for the side to move, plies uint8 until the end under optimal play,
        best_action uint8: lowest optimal action, 255 for finished positions).
    The reference's ``MinimaxResult.score`` is ``outcome * (win - plies)`` with ``win = EvalConfig.win``."""
    s, on_dev, d = _states_in(states)
    d = _dev(device, on_dev, d)
    lib = N.init(d)
    oc, pl, act = _out((s.n,), np.int8, on_dev, d), _out((s.n,), np.uint8, on_dev, d), _out((s.n,), np.uint8, on_dev, d)
    N.check(lib.qk_solve(s.ptr, s.n, oc.ptr, pl.ptr, act.ptr, d, _stream(on_dev, d)))
    return oc.obj, pl.obj, act.obj


def solve_moves(states, device: Optional[int] = None):
    """Exact value of EVERY legal move of each position: one qk_movegen_masks, one qk_apply_moves and one qk_solve over all
    children.  -> (score (N, 64) int16, optimal (N,) uint64): score[i, a] = the mover's result of action a as
    +(64 - plies) for a win in `plies` plies, -(64 - plies) for a loss (sooner wins and later losses score higher, the
    order of MinimaxEngine's W - ply, minimax.py:354-369), 0 for an illegal action; optimal[i] = mask of the actions with
    the best score.  The reference's best_move is one member of that set -- which one depends on the principal variation
    its shallower, heuristic iterations happened to leave (minimax.py:258-268), not on the game value; qk_solve reports
    the lowest member."""
    s, _, _ = _states_in(states)
    host = np.asarray(states.cpu().numpy() if _is_torch(states) else states).view(np.uint16).reshape(-1, 8)
    masks, _, _ = movegen_masks(host, device=device)
    bits = np.unpackbits(np.ascontiguousarray(masks).view(np.uint8).reshape(-1, 8), axis=1, bitorder="little")
    pi, act = np.nonzero(bits)
    score = np.zeros((len(host), 64), np.int16)
    if len(pi):
        kids = apply_moves(host[pi], act.astype(np.uint8), device=device)
        oc, pl, _ = solve(kids, device=device)
        # the child's value is its mover's: negate, one ply further from the parent
        score[pi, act] = -(np.asarray(oc, np.int16) * (64 - (np.asarray(pl, np.int16) + 1)))
    best = np.where(bits.astype(bool), score, np.int16(-32768)).max(axis=1, keepdims=True)
    opt = (bits.astype(bool) & (score == best)).astype(np.uint8)
    optimal = np.packbits(opt, axis=1, bitorder="little").view(np.uint64).ravel()
    return score, optimal


def random_roots(n: int, seed: int = SEED_DEFAULT, first: int = 0, min_plies: int = 5, span: int = 7,
                 device: Optional[int] = None, as_torch: bool = False):
    """Mid-game root generator of BASELINE config 4 (benchmarks/dataset.py:38-51 rejection rule)."""
    d = 0 if device is None else int(device)
    lib = N.init(d)
    out = _out((n, 8), np.uint16, as_torch, d)
    N.check(lib.qk_random_roots(int(seed) & 0xFFFFFFFFFFFFFFFF, int(first), n, int(min_plies), int(span), out.ptr,
                                d, _stream(as_torch, d)))
    return out.obj


# ----------------------------------------------------------------------------- perft
def perft(roots, depth: int, mult=None, device: Optional[int] = None) -> Dict[str, np.ndarray]:
    """Raw tree walk with SymmetryTable's counting rules (game_stats.py:384-485).
    -> uint64[depth+1] arrays nodes, wins_p0, wins_p1, blocked_p0_loses, blocked_p1_loses."""
    s, on_dev, d = _states_in(roots)
    d = _dev(device, on_dev, d)
    lib = N.init(d)
    m = _aux_in(mult, np.uint64, on_dev, d)
    if mult is not None and m.n != s.n:
        raise ValueError("one multiplicity per root required")
    arrs = [np.zeros(depth + 1, np.uint64) for _ in range(5)]
    N.check(lib.qk_perft(s.ptr, m.ptr, s.n, int(depth), *[a.ctypes.data for a in arrs], d, _stream(on_dev, d)))
    return dict(zip(["nodes", "wins_p0", "wins_p1", "blocked_p0_loses", "blocked_p1_loses"], arrs))


def perft_probe(roots, device: Optional[int] = None):
    """Cost probe for the static partitioning of an enumeration (qk_perft_probe): number of move sequences of
    length 2 below every root -> uint32 per root."""
    s, on_dev, d = _states_in(roots)
    d = _dev(device, on_dev, d)
    lib = N.init(d)
    out = _out((s.n,), np.uint32, on_dev, d)
    N.check(lib.qk_perft_probe(s.ptr, s.n, out.ptr, d, _stream(on_dev, d)))
    return out.obj


def symmetry_table_rows(p: Dict[str, np.ndarray]):
    """Re-express perft counters from the EMPTY board in the reference's table convention
    (GAME_TREE_ANALYSIS.md:4-11): row d = [total_legal_moves, p0_wins, p1_wins, ongoing], where a
    blocked mover at depth d-1 is booked as a win of the other player at depth d
    (game_stats.py:397-398,426-443) and 'ongoing' = non-won positions at depth d."""
    rows = []
    for d in range(1, len(p["nodes"])):
        w0 = int(p["wins_p0"][d]) + int(p["blocked_p1_loses"][d - 1])
        w1 = int(p["wins_p1"][d]) + int(p["blocked_p0_loses"][d - 1])
        ongoing = int(p["nodes"][d]) - int(p["wins_p0"][d]) - int(p["wins_p1"][d])
        rows.append([int(p["nodes"][d]), w0, w1, ongoing])
    return rows


def int_issue_peak(device: int = 0):
    """Measured integer issue rates (lane-ops/s): pure LOP3 stream, LOP3+IMAD mix."""
    lib = N.init(device)
    a, m = ctypes.c_double(), ctypes.c_double()
    N.check(lib.qk_int_issue_peak(device, ctypes.byref(a), ctypes.byref(m)))
    return {"alu": a.value, "mixed": m.value}


def pipe_rates(device: int = 0):
    """Measured issue rate (lane-ops/s) per instruction class; see include/quantik_b200.h qk_pipe_rates."""
    lib = N.init(device)
    out = (ctypes.c_double * 8)()
    N.check(lib.qk_pipe_rates(device, out))
    names = ["lop3", "imad", "imad_hi", "imad_wide", "popc", "lop3_imad", "lop3_imad_hi", "shf_lop3"]
    return dict(zip(names, [float(x) for x in out]))
