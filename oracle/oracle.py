"""ctypes front-end of the C oracle (``oracle/quantik_oracle.c``).

TEST INFRASTRUCTURE ONLY -- see the header of ``quantik_oracle.c``.  Parity
status: PINNED against golden vectors generated from the reference
(``oracle/gen_golden.py`` -> ``tests/golden/``).

Every function takes / returns numpy arrays in the reference's batch wire
layout (N x 8 little-endian uint16, ``storage/compact_state.py:125-147`` of the
reference).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "quantik_oracle.c")
_LIB = os.path.join(_HERE, "_build", "libquantik_oracle.so")

SEED_DEFAULT = 0x5155414E54494B31  # "QUANTIK1"


def build(force: bool = False) -> str:
    """Compile the C oracle with gcc (a few hundred ms)."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_LIB), exist_ok=True)
        subprocess.check_call(
            ["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-fvisibility=hidden",
             "-o", _LIB, _SRC])
    return _LIB


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        L = _lib
        u16p = ctypes.POINTER(ctypes.c_uint16)
        L.qo_validate.argtypes = [u16p, ctypes.POINTER(ctypes.c_int)]
        L.qo_validate.restype = ctypes.c_int
        L.qo_permute16.argtypes = [ctypes.c_uint16, ctypes.c_int]
        L.qo_permute16.restype = ctypes.c_uint16
        L.qo_symmetry_table.restype = ctypes.c_long
        L.qo_num_threads.restype = ctypes.c_int
    return _lib


def _states(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint16)
    if a.ndim == 1:
        a = a.reshape(-1, 8)
    assert a.ndim == 2 and a.shape[1] == 8
    return a


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def num_threads() -> int:
    return int(lib().qo_num_threads())


def set_num_threads(n: int) -> None:
    lib().qo_set_num_threads(ctypes.c_int(n))


def validate(states) -> Tuple[np.ndarray, np.ndarray]:
    """-> (player or 255, ValidationResult code) per state."""
    s = _states(states)
    player = np.empty(len(s), np.uint8)
    code = np.empty(len(s), np.uint8)
    p = ctypes.c_int()
    for i in range(len(s)):
        code[i] = lib().qo_validate(s[i].ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)), ctypes.byref(p))
        player[i] = 255 if p.value < 0 else p.value
    return player, code


def movegen_masks(states) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    s = _states(states)
    mask = np.empty(len(s), np.uint64)
    to_move = np.empty(len(s), np.uint8)
    status = np.empty(len(s), np.uint8)
    lib().qo_movegen_batch(_p(s), ctypes.c_size_t(len(s)), _p(mask), _p(to_move), _p(status))
    return mask, to_move, status


def win_status(states) -> np.ndarray:
    s = _states(states)
    out = np.empty(len(s), np.uint8)
    lib().qo_win_batch(_p(s), ctypes.c_size_t(len(s)), _p(out))
    return out


def apply_moves(states, actions) -> np.ndarray:
    s = _states(states)
    a = np.ascontiguousarray(actions, dtype=np.uint8)
    out = np.empty_like(s)
    lib().qo_apply_batch(_p(s), _p(a), ctypes.c_size_t(len(s)), _p(out))
    return out


def canonical(states, color_swap: bool = False, want_transform: bool = False):
    s = _states(states)
    out = np.empty_like(s)
    tr = np.empty((len(s), 7), np.int32) if want_transform else None
    lib().qo_canonical_batch(_p(s), ctypes.c_size_t(len(s)), ctypes.c_int(int(color_swap)), _p(out), _p(tr))
    return (out, tr) if want_transform else out


def orbit_size(states) -> np.ndarray:
    s = _states(states)
    out = np.empty(len(s), np.uint8)
    lib().qo_orbit_batch(_p(s), ctypes.c_size_t(len(s)), _p(out))
    return out


def apply_symmetry(state, d4: int, color_swap: bool, perm) -> np.ndarray:
    s = _states(state)
    out = np.empty(8, np.uint16)
    pa = (ctypes.c_int * 4)(*perm)
    lib().qo_apply_symmetry(_p(s), ctypes.c_int(d4), ctypes.c_int(int(color_swap)), pa, _p(out))
    return out


def perm(index: int):
    pa = (ctypes.c_int * 4)()
    lib().qo_perm(ctypes.c_int(index), pa)
    return tuple(pa)


def apply_symmetry_to_move(player, shape, pos, d4, color_swap, perm_):
    pa = (ctypes.c_int * 4)(*perm_)
    out = (ctypes.c_int * 3)()
    lib().qo_apply_symmetry_to_move(player, shape, pos, d4, int(color_swap), pa, out)
    return tuple(out)


def philox4x32_10(ctr, key) -> np.ndarray:
    c = np.ascontiguousarray(ctr, dtype=np.uint32)
    k = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.empty(4, np.uint32)
    lib().qo_philox4x32_10(_p(c), _p(k), _p(out))
    return out


def playouts(roots, rollouts: int, seed: int = SEED_DEFAULT, first_rollout: int = 0,
             root_index_base: int = 0, max_plies: int = -1, per_playout: bool = False):
    """-> dict(wins_p0, wins_p1, draws[, outcome (n_roots, rollouts) int8, plies uint8])."""
    s = _states(roots)
    n = len(s)
    w0 = np.zeros(n, np.uint32)
    w1 = np.zeros(n, np.uint32)
    dr = np.zeros(n, np.uint32)
    oc = np.zeros((n, rollouts), np.int8) if per_playout else None
    pl = np.zeros((n, rollouts), np.uint8) if per_playout else None
    lib().qo_playouts(_p(s), ctypes.c_size_t(n), ctypes.c_uint32(rollouts), ctypes.c_uint64(seed),
                      ctypes.c_uint64(first_rollout), ctypes.c_uint32(root_index_base),
                      ctypes.c_int(max_plies), _p(w0), _p(w1), _p(dr), _p(oc), _p(pl))
    out = {"wins_p0": w0, "wins_p1": w1, "draws": dr}
    if per_playout:
        out["outcome"] = oc
        out["plies"] = pl
    return out


def make_roots(n: int, seed: int = SEED_DEFAULT, first: int = 0, min_plies: int = 5, span: int = 7) -> np.ndarray:
    out = np.empty((n, 8), np.uint16)
    lib().qo_make_roots(ctypes.c_uint64(seed), ctypes.c_size_t(first), ctypes.c_size_t(n),
                        ctypes.c_int(min_plies), ctypes.c_int(span), _p(out))
    return out


def perft(roots, depth: int, mult=None):
    """-> dict of uint64[depth+1]: nodes, wins_p0, wins_p1, blocked_p0_loses, blocked_p1_loses."""
    s = _states(roots)
    m = None if mult is None else np.ascontiguousarray(mult, dtype=np.uint64)
    arrs = [np.zeros(depth + 1, np.uint64) for _ in range(5)]
    lib().qo_perft(_p(s), _p(m), ctypes.c_size_t(len(s)), ctypes.c_int(depth), *[_p(a) for a in arrs])
    return dict(zip(["nodes", "wins_p0", "wins_p1", "blocked_p0_loses", "blocked_p1_loses"], arrs))


def symmetry_table(max_depth: int, emit_depth: int = 0, cap: int = 0):
    """The reference's SymmetryTable: rows of (total_legal_moves, unique, p0_wins, p1_wins, ongoing)
    for depth 1..max_depth; optionally the canonical states + multiplicities of ``emit_depth``."""
    stats = np.zeros((max_depth, 5), np.uint64)
    st = np.zeros((cap, 8), np.uint16) if cap else None
    mu = np.zeros(cap, np.uint64) if cap else None
    n = lib().qo_symmetry_table(ctypes.c_int(max_depth), _p(stats), ctypes.c_int(emit_depth),
                                _p(st), _p(mu), ctypes.c_size_t(cap))
    if cap:
        return stats, st[:n], mu[:n]
    return stats


SEEDED_WEIGHTS = (100.0, -100.0, 20.0, 3.0, 2.0, 0.0)     # evaluation.py:40


def features(states, players) -> np.ndarray:
    """evaluation.features (evaluation.py:130-194) -> (N, 6) float64."""
    s = _states(states)
    pl = np.ascontiguousarray(players, dtype=np.uint8)
    out = np.zeros((len(s), 6), np.float64)
    lib().qo_features_batch(_p(s), _p(pl), ctypes.c_size_t(len(s)), _p(out))
    return out


def guided_playouts(roots, rollouts: int, seed: int = SEED_DEFAULT, epsilon: float = 0.2, weights=SEEDED_WEIGHTS,
                    first_rollout: int = 0, root_index_base: int = 0, max_plies: int = -1, per_playout: bool = False):
    """Playouts with MCTSEngine._select_rollout_move's eval-guided policy (mcts.py:345-383)."""
    s = _states(roots)
    n = len(s)
    w = np.ascontiguousarray(weights, dtype=np.float64)
    w0, w1, dr = np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(n, np.uint32)
    oc = np.zeros((n, rollouts), np.int8) if per_playout else None
    pl = np.zeros((n, rollouts), np.uint8) if per_playout else None
    lib().qo_guided_playouts(_p(s), ctypes.c_size_t(n), ctypes.c_uint32(rollouts), ctypes.c_uint64(seed),
                             ctypes.c_uint64(first_rollout), ctypes.c_uint32(root_index_base), ctypes.c_int(max_plies),
                             ctypes.c_double(epsilon), _p(w), _p(w0), _p(w1), _p(dr), _p(oc), _p(pl))
    out = {"wins_p0": w0, "wins_p1": w1, "draws": dr}
    if per_playout:
        out["outcome"], out["plies"] = oc, pl
    return out


def solve(states):
    """Exact game value of each position (minimax.py:330-453 at full depth): outcome +1 / -1 for the side to
    move, plies until the end under optimal play, lowest optimal action (255 when the position is finished)."""
    s = _states(states)
    oc, pl, act = np.zeros(len(s), np.int8), np.zeros(len(s), np.uint8), np.zeros(len(s), np.uint8)
    lib().qo_solve_batch(_p(s), ctypes.c_size_t(len(s)), _p(oc), _p(pl), _p(act))
    return oc, pl, act


def opening_book_summary(max_depth: int) -> np.ndarray:
    """rows (positions, terminal, edges) for depth 0..max_depth, as opening_book_summary.build_summary counts them."""
    out = np.zeros((max_depth + 1, 3), np.uint64)
    lib().qo_opening_book_summary(ctypes.c_int(max_depth), _p(out))
    return out


__all__ = [n for n in dir() if not n.startswith("_")]
