"""Drop-in mirror of the reference's single-state hot-path API, backed by the CUDA library.

Same names, argument meaning and error behaviour as ``quantik_core``'s
``move`` / ``game_utils`` / ``state_validator`` / ``symmetry`` / ``core`` modules, so an
engine that does ``from quantik_core.move import generate_legal_moves`` can bind these
instead.  Every rule evaluation goes through
``libquantik_b200.so``; nothing here re-implements the rules on the CPU (only the text
codec QFEN and the trivial move re-labelling of ``apply_symmetry_to_move`` are host
arithmetic, as in SURVEY 8a rows a11 / component 10).

One call = one kernel launch for one state: correct but latency-bound.  The batched
forms in :mod:`quantik_b200.batch` are the fast path.
"""
from __future__ import annotations

import itertools
import struct
from dataclasses import dataclass
from enum import IntEnum
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import batch as B

Bitboard = Tuple[int, int, int, int, int, int, int, int]   # commons.py:13
PlayerId = int

VERSION = 1                 # commons.py:51
FLAG_CANON = 1 << 1         # commons.py:52
MAX_PIECES_PER_SHAPE = 2    # commons.py:20
SHAPE_LETTERS = "ABCD"      # qfen.py:72


class ValidationError(Exception):
    """commons.py:55-58"""


class ValidationResult(IntEnum):   # state_validator.py:16-27
    OK = 0
    TURN_BALANCE_INVALID = 1
    SHAPE_COUNT_EXCEEDED = 2
    NOT_PLAYER_TURN = 3
    ILLEGAL_PLACEMENT = 4
    PIECE_OVERLAP = 5
    INVALID_POSITION = 6
    INVALID_SHAPE = 7
    INVALID_PLAYER = 8


class WinStatus(IntEnum):          # game_utils.py:32-37
    NO_WIN = 0
    PLAYER_0_WINS = 1
    PLAYER_1_WINS = 2


# ------------------------------------------------------------------ argument checks
def validate_move_parameters(player: int, shape: int, position: int) -> None:
    """game_utils.py:289-341: ValueError for a bad player / shape / position."""
    if player not in (0, 1):
        raise ValueError(f"Invalid player: {player}")
    if shape not in range(4):
        raise ValueError(f"Invalid shape: {shape}")
    if position not in range(16):
        raise ValueError(f"Invalid position: {position}")


@dataclass(frozen=True)
class Move:
    """move.py:24-39"""

    player: PlayerId
    shape: int
    position: int

    def __post_init__(self) -> None:
        validate_move_parameters(self.player, self.shape, self.position)


def _bb_tuple(bb) -> Bitboard:
    if hasattr(bb, "to_tuple"):          # CompactBitboard-like
        bb = bb.to_tuple()
    t = tuple(int(x) for x in bb)
    if len(t) != 8:
        raise ValueError("Invalid bitboard data")                       # core.py:30-31
    for v in t:
        if v < 0 or v > 0xFFFF:
            raise ValueError("Bitboard values must be between 0 and 65535")   # bitboard_compact.py:56-58
    return t   # type: ignore[return-value]


def _arr(bb) -> np.ndarray:
    return np.array([_bb_tuple(bb)], dtype=np.uint16)


# ------------------------------------------------------------------ state_validator
def validate_game_state(bb, raise_on_error: bool = False) -> Tuple[Optional[PlayerId], ValidationResult]:
    """state_validator.py:171-196"""
    _, to_move, status = B.movegen_masks(_arr(bb))
    result = ValidationResult(int(status[0]))
    player = int(to_move[0]) if result == ValidationResult.OK else None
    if raise_on_error and result != ValidationResult.OK:
        raise ValidationError(f"Invalid game state: {result}")
    return player, result


# ------------------------------------------------------------------ move
def _moves_from_mask(mask: int, player: int) -> Dict[int, List[Move]]:
    return {s: [Move(player, s, p) for p in range(16) if (mask >> (s * 16 + p)) & 1] for s in range(4)}


def generate_legal_moves(bb, player_id: Optional[PlayerId] = None) -> Tuple[PlayerId, Dict[int, List[Move]]]:
    """move.py:199-268: (current_player, {shape: [Move]}); (0, {}) for an invalid state,
    (current_player, {}) when ``player_id`` is not the side to move."""
    mask, to_move, status = B.movegen_masks(_arr(bb))
    if int(status[0]) != 0:
        return 0, {}
    current = int(to_move[0])
    if player_id is not None and player_id != current:
        return current, {}
    return current, _moves_from_mask(int(mask[0]), current)


def generate_legal_moves_list(bb, player_id: Optional[PlayerId] = None) -> List[Move]:
    """move.py:271-291"""
    _, by_shape = generate_legal_moves(bb, player_id)
    return [m for ms in by_shape.values() for m in ms]


def legal_action_mask(bb) -> int:
    """64-bit action mask, bit = shape*16+pos (api_portability_report.py:79-86)."""
    return int(B.movegen_masks(_arr(bb))[0][0])


def apply_move(bb, move: Move):
    """move.py:135-166: the bit of ``move`` OR-ed into word player*4+shape; no legality check.
    Returns the same container type it was given (tuple, or an object with ``from_tuple``)."""
    action = 0x80 | (move.player << 6) | (move.shape * 16 + move.position)
    out = B.apply_moves(_arr(bb), np.array([action], np.uint8))
    t = tuple(int(x) for x in out[0])
    if hasattr(bb, "to_tuple") and hasattr(type(bb), "from_tuple"):
        return type(bb).from_tuple(t)
    return t


class MoveValidationResult:
    """move.py:42-55"""

    def __init__(self, is_valid: bool, error: Optional[ValidationResult] = None, new_bb=None):
        self.is_valid = is_valid
        self.error = error
        self.new_bb = new_bb


def validate_move(bb, move: Move) -> MoveValidationResult:
    """move.py:58-132: turn check, empty-square check, validity of the resulting position."""
    t = _bb_tuple(bb)
    current, result = validate_game_state(t)
    if result != ValidationResult.OK:
        return MoveValidationResult(False, result)
    if current != move.player:
        return MoveValidationResult(False, ValidationResult.NOT_PLAYER_TURN)
    if any(v >> move.position & 1 for v in t):
        return MoveValidationResult(False, ValidationResult.PIECE_OVERLAP)
    new_bb = apply_move(bb, move)
    _, new_result = validate_game_state(new_bb)
    if new_result != ValidationResult.OK:
        return MoveValidationResult(False, new_result)
    return MoveValidationResult(True, None, new_bb)


# ------------------------------------------------------------------ game_utils
def check_game_winner(bb) -> WinStatus:
    """game_utils.py:161-189"""
    return WinStatus(int(B.win_status(_arr(bb))[0]))


def has_winning_line(bb) -> bool:
    """game_utils.py:123-158"""
    return int(B.win_status(_arr(bb))[0]) != 0


# ------------------------------------------------------------------ qfen (host text codec, qfen.py:89-168)
def bb_to_qfen(bb) -> str:
    t = _bb_tuple(bb)
    rows = []
    for r in range(4):
        row = []
        for c in range(4):
            i = r * 4 + c
            ch = "."
            for color in (0, 1):
                for s in range(4):
                    if (t[color * 4 + s] >> i) & 1:
                        ch = SHAPE_LETTERS[s] if color == 0 else SHAPE_LETTERS[s].lower()
            row.append(ch)
        rows.append("".join(row))
    return "/".join(rows)


def bb_from_qfen(qfen: str, validate: bool = False) -> Bitboard:
    parts = [p.strip() for p in qfen.replace(" ", "").split("/")]
    if len(parts) != 4 or any(len(p) != 4 for p in parts):
        raise ValueError("QFEN must be 4 ranks of 4 chars separated by '/'")
    bb = [0] * 8
    for r in range(4):
        for c in range(4):
            ch = parts[r][c]
            if ch == ".":
                continue
            if ch.upper() not in SHAPE_LETTERS:
                raise ValueError(f"Invalid character '{ch}' in QFEN. Must be A,B,C,D (uppercase/lowercase) or '.'")
            bb[(0 if ch.isupper() else 4) + SHAPE_LETTERS.index(ch.upper())] |= 1 << (r * 4 + c)
    t = tuple(bb)
    if validate:
        _, result = validate_game_state(t)
        if result != ValidationResult.OK:
            raise ValidationError(f"Invalid qfen: {qfen}. {str(result)}")
    return t   # type: ignore[return-value]


# ------------------------------------------------------------------ symmetry
D4_NAMES = ("id", "rot90", "rot180", "rot270", "reflV", "reflH", "reflD", "reflAD")   # symmetry.py:101-110
_D4_FN = (lambda r, c: (r, c), lambda r, c: (c, 3 - r), lambda r, c: (3 - r, 3 - c), lambda r, c: (3 - c, r),
          lambda r, c: (r, 3 - c), lambda r, c: (3 - r, c), lambda r, c: (c, r), lambda r, c: (3 - c, 3 - r))
_D4_INVERSE = (0, 3, 2, 1, 4, 5, 6, 7)                                                # symmetry.py:116-125


@dataclass(frozen=True)
class SymmetryTransform:
    """symmetry.py:45-85"""

    d4_index: int
    color_swap: bool
    shape_perm: Tuple[int, ...]

    def __post_init__(self) -> None:
        if not 0 <= self.d4_index < 8:
            raise ValueError(f"D4 index must be 0-7, got {self.d4_index}")
        if len(self.shape_perm) != 4 or set(self.shape_perm) != set(range(4)):
            raise ValueError(f"Shape perm must be a permutation of [0,1,2,3], got {self.shape_perm}")

    def inverse(self) -> "SymmetryTransform":
        inv = [0] * 4
        for i, j in enumerate(self.shape_perm):
            inv[j] = i
        return SymmetryTransform(_D4_INVERSE[int(self.d4_index)], self.color_swap, tuple(inv))

    def code(self) -> int:
        return B.encode_transform(int(self.d4_index), self.color_swap, self.shape_perm)


class SymmetryHandler:
    """symmetry.py:88-480 (the parts on the hot path)."""

    ALL_SHAPE_PERMS = list(itertools.permutations(range(4)))

    @classmethod
    def get_d4_inverse(cls, d4_index: int) -> int:
        return _D4_INVERSE[int(d4_index)]

    @classmethod
    def permute16(cls, mask: int, d4_index: int) -> int:
        """symmetry.py:237-250"""
        out = B.apply_symmetry(np.array([[mask & 0xFFFF, 0, 0, 0, 0, 0, 0, 0]], np.uint16),
                               np.array([B.encode_transform(int(d4_index), False, (0, 1, 2, 3))], np.uint16))
        return int(out[0][0])

    @classmethod
    def apply_symmetry(cls, bb, transform: SymmetryTransform) -> Bitboard:
        """symmetry.py:252-287"""
        out = B.apply_symmetry(_arr(bb), np.array([transform.code()], np.uint16))
        return tuple(int(x) for x in out[0])   # type: ignore[return-value]

    @classmethod
    def find_canonical_form(cls, bb, color_swap: bool = False) -> Tuple[Bitboard, SymmetryTransform]:
        """symmetry.py:289-354"""
        pay, tr = B.canonical_payloads(_arr(bb), color_swap=color_swap, want_transform=True)
        d4, cs, perm = B.decode_transform(int(tr[0]))
        return tuple(int(x) for x in pay[0]), SymmetryTransform(d4, cs, perm)   # type: ignore[return-value]

    @classmethod
    def count_orbit_size(cls, bb) -> int:
        """symmetry.py:356-396"""
        return int(B.orbit_sizes(_arr(bb))[0])

    @classmethod
    def get_canonical_payload(cls, bb) -> bytes:
        """symmetry.py:398-410"""
        return np.ascontiguousarray(B.canonical_payloads(_arr(bb))[0]).astype("<u2").tobytes()

    @classmethod
    def get_canonical_key(cls, bb) -> bytes:
        """symmetry.py:412-423"""
        return bytes([VERSION, FLAG_CANON]) + cls.get_canonical_payload(bb)

    @classmethod
    def apply_symmetry_to_move(cls, move: Move, transform: SymmetryTransform) -> Move:
        """symmetry.py:425-464"""
        r2, c2 = _D4_FN[int(transform.d4_index)](move.position // 4, move.position % 4)
        player = 1 - move.player if transform.color_swap else move.player
        return Move(player=player, shape=tuple(transform.shape_perm).index(move.shape), position=r2 * 4 + c2)

    @classmethod
    def get_qfen_canonical_form(cls, qfen: str) -> str:
        """symmetry.py:466-480"""
        return bb_to_qfen(cls.find_canonical_form(bb_from_qfen(qfen))[0])


# ------------------------------------------------------------------ core.State
class State:
    """core.py:12-236 (binary core, QFEN, canonicalisation; CBOR left to the reference)."""

    __slots__ = ("_bb",)

    def __init__(self, bb) -> None:
        object.__setattr__(self, "_bb", _bb_tuple(bb))

    def __setattr__(self, *_):   # frozen like the reference dataclass
        raise AttributeError("State is immutable")

    def __eq__(self, other) -> bool:
        return isinstance(other, State) and other._bb == self._bb

    def __hash__(self) -> int:
        return hash(self._bb)

    def __repr__(self) -> str:
        return f"State({self._bb})"

    @property
    def bb(self) -> Bitboard:
        return self._bb

    @staticmethod
    def empty() -> "State":
        return State((0,) * 8)

    def pack(self, flags: int = 0) -> bytes:
        return struct.pack("<BB8H", VERSION, flags, *self._bb)          # core.py:52-54

    @staticmethod
    def unpack(data: bytes) -> "State":
        if len(data) < 18:
            raise ValueError("Buffer too small for v1 core (18 bytes).")   # core.py:60-61
        ver, _flags, *rest = struct.unpack("<BB8H", data[:18])
        if ver != VERSION:
            raise ValueError(f"Unsupported version {ver}")                  # core.py:63-64
        return State(tuple(int(x) & 0xFFFF for x in rest))

    def to_qfen(self) -> str:
        return bb_to_qfen(self._bb)

    @staticmethod
    def from_qfen(qfen: str, validate: bool = False) -> "State":
        return State(bb_from_qfen(qfen, validate))

    def canonical_payload(self) -> bytes:
        return SymmetryHandler.get_canonical_payload(self._bb)             # core.py:162-170

    def canonical_key(self) -> bytes:
        return SymmetryHandler.get_canonical_key(self._bb)                 # core.py:171-180

    def symmetry_count(self) -> int:
        return SymmetryHandler.count_orbit_size(self._bb)                  # core.py:182-192

    def get_occupied_bb(self) -> int:
        o = 0
        for v in self._bb:
            o |= v
        return o

    # CBOR wrappers (core.py:194-228): { "v":1, "canon":bool, "bb": h'16 bytes', ? "mc":uint, ? "meta":{...} }
    def to_cbor(self, canon: bool = False, mc: Optional[int] = None, meta: Optional[Dict] = None) -> bytes:
        try:
            import cbor2
        except ImportError:
            raise RuntimeError("Please install cbor2 (pip install cbor2)")
        m = {"v": VERSION, "canon": bool(canon), "bb": struct.pack("<8H", *self._bb)}
        if mc is not None:
            m["mc"] = int(mc)
        if meta:
            m["meta"] = meta
        return cbor2.dumps(m)

    @staticmethod
    def from_cbor(data: bytes) -> "State":
        try:
            import cbor2
        except ImportError:
            raise RuntimeError("Please install cbor2 (pip install cbor2)")
        m = cbor2.loads(data)
        if m.get("v") != VERSION:
            raise ValueError("Unsupported CBOR version")
        bb = m.get("bb")
        if not isinstance(bb, (bytes, bytearray)) or len(bb) != 16:
            raise ValueError("CBOR field 'bb' must be 16 bytes")
        return State(tuple(int(x) & 0xFFFF for x in struct.unpack("<8H", bb)))


# ------------------------------------------------------------------ conformance record
def portability_case_report(case_id: str, qfen: str, shape: int, position: int) -> dict:
    """The per-case record of api_portability_report._case_report (api_portability_report.py:152-191),
    produced from the CUDA path: what the reference's cross-stack check compares field by field."""
    bb = bb_from_qfen(qfen)
    a = _arr(bb)
    mask, to_move, status = B.movegen_masks(a)
    if int(status[0]) != 0:
        raise ValueError(f"{case_id}: invalid game state {ValidationResult(int(status[0])).name}")
    side = int(to_move[0])
    mask = int(mask[0])
    indices = [i for i in range(64) if (mask >> i) & 1]
    win = int(B.win_status(a)[0])
    winner = {0: "none", 1: "player0", 2: "player1"}[win]
    if winner != "none":
        terminal = True
    elif indices:
        terminal = False
    else:
        terminal, winner = True, ("player1" if side == 0 else "player0")
    canonical_bb = B.canonical_payloads(a)[0]
    move = Move(side, shape, position)
    v = validate_move(bb, move)
    return {
        "case_id": case_id, "qfen": bb_to_qfen(bb), "bitboards": list(bb), "side_to_move": side,
        "canonical_qfen": bb_to_qfen(tuple(int(x) for x in canonical_bb)),
        "canonical_key": (bytes([VERSION, FLAG_CANON]) + np.ascontiguousarray(canonical_bb).astype("<u2").tobytes()).hex(),
        "orbit_size": int(B.orbit_sizes(a)[0]),
        "legal_action_mask": f"0x{mask:016x}", "legal_action_indices": indices,
        "terminal": terminal, "winner": winner,
        "move": {"shape": shape, "position": position, "action_index": shape * 16 + position,
                 "is_legal": v.is_valid, "after_qfen": bb_to_qfen(v.new_bb) if v.is_valid else None},
    }


def portability_report(cases, contracts_release: str = "", contract_ids=None) -> dict:
    """The WHOLE ``api-portability-report.v1`` document of api_portability_report.build_report
    (api_portability_report.py:194-226) for a list of fixture cases ``{"case_id", "qfen", "move": {"shape", "position"}}``
    (the ``game_state_cases`` of the contracts fixture), every primitive evaluated for all cases in ONE launch:
    movegen, winner, canonical form, orbit size, the case's move applied.  The fixture / manifest files themselves are
    the caller's (they live in the contracts repository, not in quantik-core): pass their release and contract ids."""
    if not isinstance(cases, list) or not cases:
        raise ValueError("game_state_cases must be a non-empty list")
    parsed = []
    for index, case in enumerate(cases):
        if not isinstance(case, dict):
            raise ValueError(f"game_state_cases[{index}] must be an object")
        case_id, qfen, mv = case.get("case_id"), case.get("qfen"), case.get("move")
        if not isinstance(case_id, str) or not case_id:
            raise ValueError("game_state_case.case_id must be a non-empty string")
        if not isinstance(qfen, str):
            raise ValueError(f"{case_id}: qfen must be a string")
        if not isinstance(mv, dict):
            raise ValueError(f"{case_id}: move must be an object")
        try:
            bb = bb_from_qfen(qfen)
        except ValueError as exc:
            raise ValueError(f"{case_id}: qfen parse failed: {exc}") from exc
        vals = []
        for key, upper in (("shape", 3), ("position", 15)):
            value = mv.get(key)
            if not isinstance(value, int) or isinstance(value, bool):
                raise ValueError(f"move.{key} must be an integer")
            if value < 0 or value > upper:
                raise ValueError(f"move.{key} must be between 0 and {upper}")
            vals.append(value)
        parsed.append((case_id, bb, vals[0], vals[1]))
    boards = np.array([p[1] for p in parsed], np.uint16).reshape(-1, 8)
    masks, movers, status = B.movegen_masks(boards)
    for (case_id, *_), st in zip(parsed, status):
        if int(st) != 0:
            raise ValueError(f"{case_id}: invalid game state {ValidationResult(int(st)).name}")
    wins = B.win_status(boards)
    canon = B.canonical_payloads(boards)
    orbits = B.orbit_sizes(boards)
    actions = np.array([p[2] * 16 + p[3] for p in parsed], np.uint8)
    # validate_move (move.py:135-167): the fixture move is made by the side to move; legal = its bit of the mask,
    # except on a finished board, where generate_legal_moves still lists moves but nobody checks the winner
    legal = ((masks >> actions.astype(np.uint64)) & np.uint64(1)).astype(bool)
    after = B.apply_moves(boards, actions)
    out = []
    for i, (case_id, bb, shape, position) in enumerate(parsed):
        mask = int(masks[i])
        indices = [a for a in range(64) if (mask >> a) & 1]
        winner = {0: "none", 1: "player0", 2: "player1"}[int(wins[i])]
        side = int(movers[i])
        if winner != "none":
            terminal = True
        elif indices:
            terminal = False
        else:
            terminal, winner = True, ("player1" if side == 0 else "player0")
        out.append({
            "case_id": case_id, "qfen": bb_to_qfen(bb), "bitboards": list(bb), "side_to_move": side,
            "canonical_qfen": bb_to_qfen(tuple(int(x) for x in canon[i])),
            "canonical_key": (bytes([VERSION, FLAG_CANON]) + np.ascontiguousarray(canon[i]).astype("<u2").tobytes()).hex(),
            "orbit_size": int(orbits[i]),
            "legal_action_mask": f"0x{mask:016x}", "legal_action_indices": indices,
            "terminal": terminal, "winner": winner,
            "move": {"shape": shape, "position": position, "action_index": shape * 16 + position, "is_legal": bool(legal[i]),
                     "after_qfen": bb_to_qfen(tuple(int(x) for x in after[i])) if legal[i] else None},
        })
    from . import __version__
    return {"schema": "api-portability-report.v1", "contracts_release": contracts_release,
            "implementation": {"language": "cuda", "package": "quantik-core-py_b200", "version": __version__},
            "contract_ids": dict(contract_ids or {}), "cases": sorted(out, key=lambda item: item["case_id"])}

