#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ by importing THE REFERENCE ITSELF.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python oracle/gen_golden.py            # ~3 min (SymmetryTable depth 5 + 20k reference playouts)

TEST INFRASTRUCTURE ONLY.  Nothing here is imported by the product.  The vectors
it writes are what pins the C oracle (tests/test_oracle_golden.py) and, through
the oracle and directly, the CUDA kernels (tests/test_gpu_*.py).

Everything below calls the unmodified reference functions:
  quantik_core.move.generate_legal_moves / apply_move          (move.py:135-268)
  quantik_core.game_utils.check_game_winner / has_winning_line (game_utils.py:123-189)
  quantik_core.state_validator._validate_game_state_single_pass(state_validator.py:126-168)
  quantik_core.symmetry.SymmetryHandler.*                      (symmetry.py:252-464)
  quantik_core.game_stats.SymmetryTable                        (game_stats.py:276-505)
The only non-reference ingredient is the Philox4x32-10 counter stream that
replaces ``random.choice`` so that playouts are reproducible on the GPU
(DESIGN.md "Random stream"); it is checked against the Random123 known answers.
"""
from __future__ import annotations

import hashlib
import json
import os
import random
import struct
import sys
import time

import numpy as np

REF = os.environ.get("QUANTIK_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(REF, "src"))
sys.path.insert(0, os.path.join(REF, "tests"))

from quantik_core import State  # noqa: E402
from quantik_core.game_stats import SymmetryTable  # noqa: E402
from quantik_core.game_utils import WinStatus, check_game_winner, has_winning_line  # noqa: E402
from quantik_core.move import Move, apply_move, generate_legal_moves, validate_move  # noqa: E402
from quantik_core.qfen import bb_from_qfen, bb_to_qfen  # noqa: E402
from quantik_core.state_validator import _validate_game_state_single_pass  # noqa: E402
from quantik_core.symmetry import SymmetryHandler, SymmetryTransform  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
SEED = 0x5155414E54494B31  # "QUANTIK1"
M32 = 0xFFFFFFFF


# --------------------------------------------------------------------------- Philox4x32-10
def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c3 ^ k1) & M32, p0 & M32
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return (c0, c1, c2, c3)


assert philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
assert philox4x32_10((M32,) * 4, (M32,) * 2) == (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)
assert philox4x32_10((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == (
    0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)


def stream_word(seed, rollout, root_index, ply):
    out = philox4x32_10((rollout & M32, (rollout >> 32) & M32, ply // 4, root_index),
                        (seed & M32, (seed >> 32) & M32))
    return out[ply % 4]


# --------------------------------------------------------------------------- reference helpers
def ref_all_moves(bb):
    player, by_shape = generate_legal_moves(bb)
    return player, [m for ms in by_shape.values() for m in ms]


def ref_mask(bb):
    player, moves = ref_all_moves(bb)
    mask = 0
    for m in moves:
        mask |= 1 << (m.shape * 16 + m.position)
    # reference order must be ascending action index (move.py:247,262-266)
    idx = [m.shape * 16 + m.position for m in moves]
    assert idx == sorted(idx)
    return player, mask


def ref_playout(root, seed, rollout, root_index, max_plies=-1):
    """Reference rules (beam_search.py:624-645 / mcts.py:385-437) + Philox picks."""
    bb = tuple(root)
    ply = 0
    while True:
        if max_plies >= 0 and ply >= max_plies:
            return 0, ply
        w = check_game_winner(bb)
        if w != WinStatus.NO_WIN:
            return (1 if w == WinStatus.PLAYER_0_WINS else -1), ply
        player, moves = ref_all_moves(bb)
        if not moves:
            return (-1 if player == 0 else 1), ply
        r = stream_word(seed, rollout, root_index, ply)
        bb = apply_move(bb, moves[(r * len(moves)) >> 32])
        ply += 1


def ref_validate(bb):
    player, code = _validate_game_state_single_pass(tuple(bb))
    return (255 if player is None else int(player)), int(code)


def describe(bbs):
    """All reference outputs for a list of bitboards."""
    n = len(bbs)
    out = {
        "states": np.array(bbs, np.uint16),
        "v_player": np.zeros(n, np.uint8), "v_code": np.zeros(n, np.uint8),
        "mask": np.zeros(n, np.uint64), "to_move": np.zeros(n, np.uint8),
        "win": np.zeros(n, np.uint8), "has_line": np.zeros(n, np.uint8),
        "canon192": np.zeros((n, 8), np.uint16), "tr192": np.zeros((n, 6), np.int8),
        "canon384": np.zeros((n, 8), np.uint16), "tr384": np.zeros((n, 6), np.int8),
        "orbit": np.zeros(n, np.uint8),
    }
    for i, bb in enumerate(bbs):
        bb = tuple(int(x) for x in bb)
        out["v_player"][i], out["v_code"][i] = ref_validate(bb)
        p, m = ref_mask(bb)
        out["mask"][i], out["to_move"][i] = m, p
        out["win"][i] = int(check_game_winner(bb))
        out["has_line"][i] = int(has_winning_line(bb))
        c, t = SymmetryHandler.find_canonical_form(bb)
        out["canon192"][i] = c
        out["tr192"][i] = [int(t.d4_index), int(t.color_swap), *t.shape_perm]
        c, t = SymmetryHandler.find_canonical_form(bb, color_swap=True)
        out["canon384"][i] = c
        out["tr384"][i] = [int(t.d4_index), int(t.color_swap), *t.shape_perm]
        out["orbit"][i] = SymmetryHandler.count_orbit_size(bb)
    return out


def _mt_run(seed, n=10000):
    """n playouts of the reference's beam _rollout loop with its own RNG (random.choice)."""
    mt = random.Random(seed)
    w0 = 0
    for _ in range(n):
        bb = (0,) * 8
        while True:
            w = check_game_winner(bb)
            if w != WinStatus.NO_WIN:
                w0 += w == WinStatus.PLAYER_0_WINS
                break
            player, moves = ref_all_moves(bb)
            if not moves:
                w0 += player == 1
                break
            bb = apply_move(bb, mt.choice(moves))
    return int(w0)


def main():
    os.makedirs(OUT, exist_ok=True)
    t0 = time.time()
    rng = random.Random(20261019)

    # ---- 1. the reference's own tagged QFEN fixtures (tests/fixtures.py:90-697)
    from fixtures import UnifiedTestCases
    cases = []
    for tc in UnifiedTestCases.all_cases():
        bb = bb_from_qfen(tc.qfen)
        d = describe([bb])
        cases.append({
            "qfen": tc.qfen, "name": tc.name, "category": tc.category,
            "expected_player": tc.expected_player, "winner": tc.winner, "is_terminal": tc.is_terminal,
            "bb": list(bb), "v_player": int(d["v_player"][0]), "v_code": int(d["v_code"][0]),
            "mask": f"0x{int(d['mask'][0]):016x}", "to_move": int(d["to_move"][0]), "win": int(d["win"][0]),
            "canonical_key": (bytes([1, 2]) + struct.pack("<8H", *d["canon192"][0].tolist())).hex(),
            "canon384": d["canon384"][0].tolist(), "orbit": int(d["orbit"][0]),
        })
    json.dump({"source": "tests/fixtures.py UnifiedTestCases.all_cases() + reference outputs", "cases": cases},
              open(os.path.join(OUT, "fixture_cases.json"), "w"), indent=1)
    print(f"fixture_cases: {len(cases)} cases  [{time.time()-t0:.1f}s]")

    # ---- 2. states met in random reference games + corrupted (invalid) states
    states = []
    for g in range(400):
        bb = (0,) * 8
        while True:
            states.append(bb)
            if check_game_winner(bb) != WinStatus.NO_WIN:
                break
            _, moves = ref_all_moves(bb)
            if not moves:
                break
            bb = apply_move(bb, rng.choice(moves))
    states = list(dict.fromkeys(states))
    rng.shuffle(states)
    valid = states[:3000]
    invalid = []
    for bb in valid[:600]:
        b = list(bb)
        for _ in range(rng.choice([1, 1, 2, 3])):
            b[rng.randrange(8)] ^= 1 << rng.randrange(16)
        invalid.append(tuple(b))
    for _ in range(200):  # sparse garbage incl. >2 pieces per word, overlaps, bad balance
        invalid.append(tuple(sum(1 << rng.randrange(16) for _ in range(rng.choice([0, 0, 1, 1, 2, 3])))
                             & 0xFFFF for _ in range(8)))
    invalid += [(0xFFFF,) * 8, (0xFFFF, 0, 0, 0, 0, 0, 0, 0), (1, 1, 0, 0, 0, 0, 0, 0), (0, 0, 0, 0, 1, 0, 0, 0),
                (3, 0, 0, 0, 0, 0, 0, 0), (1, 2, 4, 8, 0, 0, 0, 0)]
    d = describe(valid + invalid)
    d["n_valid"] = np.array(len(valid))
    np.savez_compressed(os.path.join(OUT, "states_reference.npz"), **d)
    print(f"states_reference: {len(valid)} reachable + {len(invalid)} corrupted; codes "
          f"{np.bincount(d['v_code']).tolist()}  [{time.time()-t0:.1f}s]")

    # ---- 3. BASELINE config 3 input: benchmarks/depth4-canonical/sample-1000.json
    ds = json.load(open(os.path.join(REF, "benchmarks", "depth4-canonical", "sample-1000.json")))
    bbs = [bb_from_qfen(p["qfen"]) for p in ds["positions"]]
    keys = [State(bb).canonical_key() for bb in bbs]
    d = {
        "states": np.array(bbs, np.uint16),
        "canon192": np.array([struct.unpack("<8H", k[2:]) for k in keys], np.uint16),
        "orbit": np.array([SymmetryHandler.count_orbit_size(bb) for bb in bbs], np.uint8),
        "legal_moves": np.array([p["legal_moves"] for p in ds["positions"]], np.uint8),
        "side_to_move": np.array([p["side_to_move"] for p in ds["positions"]], np.uint8),
    }
    np.savez_compressed(os.path.join(OUT, "dataset_depth4_sample1000.npz"), **d)
    meta = {
        "source": "benchmarks/depth4-canonical/sample-1000.json", "checksum": ds["checksum"],
        "sha256_keys18": hashlib.sha256(b"".join(keys)).hexdigest(),
        "sha256_payload16": hashlib.sha256(b"".join(k[2:] for k in keys)).hexdigest(),
        "unique_keys": len(set(keys)), "sum_orbit": int(d["orbit"].astype(int).sum()),
        "first_keys": {p["id"]: k.hex() for p, k in list(zip(ds["positions"], keys))[:3]},
    }
    print("dataset:", meta)

    # ---- 4. explicit symmetry images (apply_symmetry, symmetry.py:252-287) and move transforms
    img_states, img_tr, img_out = [], [], []
    mv_in, mv_out = [], []
    for bb in bbs[:12] + valid[:12]:
        for d4 in range(8):
            for cs in (False, True):
                for perm in SymmetryHandler.ALL_SHAPE_PERMS:
                    if rng.random() > 0.25:
                        continue
                    t = SymmetryTransform(d4, cs, perm)
                    img_states.append(bb)
                    img_tr.append([d4, int(cs), *perm])
                    img_out.append(SymmetryHandler.apply_symmetry(bb, t))
                    m = Move(rng.randrange(2), rng.randrange(4), rng.randrange(16))
                    m2 = SymmetryHandler.apply_symmetry_to_move(m, t)
                    mv_in.append([m.player, m.shape, m.position])
                    mv_out.append([m2.player, m2.shape, m2.position])
    np.savez_compressed(os.path.join(OUT, "symmetry_images.npz"),
                        states=np.array(img_states, np.uint16), transform=np.array(img_tr, np.int8),
                        image=np.array(img_out, np.uint16), move_in=np.array(mv_in, np.int8),
                        move_out=np.array(mv_out, np.int8),
                        d4_maps=np.array([SymmetryHandler.build_perm(fn) for _, fn in SymmetryHandler.D4], np.uint8))
    print(f"symmetry_images: {len(img_states)}  [{time.time()-t0:.1f}s]")

    # ---- 5. SymmetryTable (the reference's perft) to depth 5 + the depth-4 canonical set
    tab = SymmetryTable()
    tab._initialize_analysis()
    rows = []
    d4_states = None
    for depth in range(1, 6):
        st = tab._process_depth(depth)
        rows.append([st.total_legal_moves, st.unique_canonical_states, st.player_0_wins, st.player_1_wins,
                     st.ongoing_games])
        if depth == 4:
            d4_states = [(s.canonical_bb, s.multiplicity) for s in tab.state_queue if s.depth == 4]
        print(f"  SymmetryTable depth {depth}: {rows[-1]}  [{time.time()-t0:.1f}s]")
    d4_states.sort(key=lambda sm: struct.pack("<8H", *sm[0]))
    np.savez_compressed(os.path.join(OUT, "depth4_canonical.npz"),
                        states=np.array([s for s, _ in d4_states], np.uint16),
                        mult=np.array([m for _, m in d4_states], np.uint64))
    meta["symmetry_table_measured"] = rows
    meta["depth4_canonical"] = {"n": len(d4_states), "sum_mult": int(sum(m for _, m in d4_states))}

    # ---- 6. playouts: reference rules driven by the shared Philox stream
    n_empty = 20000
    oc = np.zeros(n_empty, np.int8)
    pl = np.zeros(n_empty, np.uint8)
    for j in range(n_empty):
        oc[j], pl[j] = ref_playout((0,) * 8, SEED, j, 0)
    print(f"  empty-board playouts: P0 {int((oc > 0).sum())} P1 {int((oc < 0).sum())} mean plies {pl.mean():.3f}"
          f"  [{time.time()-t0:.1f}s]")
    mid = [bb for bb in valid if 5 <= sum(bin(x).count('1') for x in bb) <= 11
           and check_game_winner(bb) == WinStatus.NO_WIN and ref_all_moves(bb)[1]][:24]
    special = [bb_from_qfen("A..C/bbd./CD.A/.adB"),            # blocked mover (tests/test_board.py:265-274)
               bb_from_qfen("AbCd/..../..../...."),            # already won root
               (1, 1, 0, 0, 0, 0, 0, 0)]                       # invalid root -> (0, {}) -> P0 "loses"
    roots = mid + special
    R = 96
    moc = np.zeros((len(roots), R), np.int8)
    mpl = np.zeros((len(roots), R), np.uint8)
    for r, bb in enumerate(roots):
        for j in range(R):
            moc[r, j], mpl[r, j] = ref_playout(bb, SEED, 1000 + j, 7 + r)
    # MCTS cut-off semantics (mcts.py:412,436-437): empty root, max 16 / 12 / 5 / 0 plies
    cut = {}
    for mp in (16, 12, 5, 0):
        res = [ref_playout((0,) * 8, SEED, j, 0, mp) for j in range(3000)]
        cut[f"cut{mp}_outcome"] = np.array([a for a, _ in res], np.int8)
        cut[f"cut{mp}_plies"] = np.array([b for _, b in res], np.uint8)
    np.savez_compressed(os.path.join(OUT, "playouts_philox.npz"),
                        seed=np.array(SEED, np.uint64), empty_outcome=oc, empty_plies=pl,
                        roots=np.array(roots, np.uint16), roots_first_rollout=np.array(1000),
                        roots_index_base=np.array(7), roots_outcome=moc, roots_plies=mpl, **cut)
    print(f"playouts_philox done  [{time.time()-t0:.1f}s]")

    # ---- 7. reference's own RNG (MT19937 random.choice) tallies for the 3-sigma check
    from concurrent.futures import ProcessPoolExecutor
    mt_seeds = [12345, 1, 2, 3, 4, 5, 6, 7]
    with ProcessPoolExecutor(8) as ex:
        mt_wins = list(ex.map(_mt_run, mt_seeds))
    meta["mt19937_empty_board"] = {"n": 10000 * len(mt_seeds), "p0_wins": int(sum(mt_wins)), "per_seed": mt_wins,
                                   "seeds": mt_seeds, "note": "random.Random(seed).choice, beam _rollout loop, 10000 each"}
    meta["survey_mt19937_100k"] = {"n": 100000, "p0_wins": 48534, "note": "SURVEY.md section 6, measured with the reference"}

    # ---- 8. api-portability style records for the four cases pinned by the reference's tests
    from quantik_core.api_portability_report import _case_report
    api = []
    for cid, q, mv in [("empty", "..../..../..../....", (0, 0)), ("mixed", "Ab../..c./...D/....", (2, 12)),
                       ("single", "A.../..../..../....", (1, 5)), ("single_illegal", "A.../..../..../....", (1, 0)),
                       ("blocked", "A..C/bbd./CD.A/.adB", (0, 1))]:
        api.append(_case_report({"case_id": cid, "qfen": q, "move": {"shape": mv[0], "position": mv[1]}}))
    meta["api_portability_cases"] = api
    json.dump(meta, open(os.path.join(OUT, "reference_meta.json"), "w"), indent=1)

    # ---- 9. opening-book-summary.v1 of a full-width depth-5 book built by the reference's own generator
    #         (examples/generate_opening_book.py -> sqlite -> opening_book_summary.build_summary; ~3 min, 8 workers)
    if "--book" in sys.argv:
        import tempfile
        from pathlib import Path
        sys.path.insert(0, os.path.join(REF, "examples"))
        import generate_opening_book as GB
        from quantik_core.opening_book_summary import build_summary
        tmp = Path(tempfile.mkdtemp())
        GB.DATA_DIR = tmp
        GB.generate(max_depth=5, db_path=tmp / "book_d5.sqlite", num_workers=8)
        json.dump(build_summary(tmp / "book_d5.sqlite", 5), open(os.path.join(OUT, "opening_book_summary_d5.json"), "w"), indent=1)
    print(f"all golden fixtures written to {os.path.normpath(OUT)}  [{time.time()-t0:.1f}s]")


if __name__ == "__main__":
    main()
