"""Pin the C oracle (oracle/quantik_oracle.c) against vectors produced by the
reference itself (oracle/gen_golden.py) and against the reference's published
known answers.  CPU only."""
import hashlib
import json
import os
import struct

import numpy as np
import pytest

import oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")


def _npz(name):
    return np.load(os.path.join(G, name))


def _meta():
    return json.load(open(os.path.join(G, "reference_meta.json")))


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
        ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
        ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
         [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
    ]
    for ctr, key, want in kat:
        assert O.philox4x32_10(ctr, key).tolist() == want


def test_states_reference_all_functions():
    d = _npz("states_reference.npz")
    s = d["states"]
    pl, code = O.validate(s)
    assert np.array_equal(code, d["v_code"])
    assert np.array_equal(pl, d["v_player"])
    mask, to_move, status = O.movegen_masks(s)
    assert np.array_equal(mask, d["mask"])
    assert np.array_equal(to_move, d["to_move"])
    assert np.array_equal(status, d["v_code"])
    assert np.array_equal(O.win_status(s), d["win"])
    c, tr = O.canonical(s, False, True)
    assert np.array_equal(c, d["canon192"])
    assert np.array_equal(tr[:, [0, 1, 3, 4, 5, 6]], d["tr192"].astype(np.int32))
    c, tr = O.canonical(s, True, True)
    assert np.array_equal(c, d["canon384"])
    assert np.array_equal(tr[:, [0, 1, 3, 4, 5, 6]], d["tr384"].astype(np.int32))
    assert np.array_equal(O.orbit_size(s), d["orbit"])


def test_fixture_cases():
    cases = json.load(open(os.path.join(G, "fixture_cases.json")))["cases"]
    assert len(cases) == 44
    s = np.array([c["bb"] for c in cases], np.uint16)
    mask, to_move, status = O.movegen_masks(s)
    win = O.win_status(s)
    canon = O.canonical(s)
    orbit = O.orbit_size(s)
    for i, c in enumerate(cases):
        assert f"0x{int(mask[i]):016x}" == c["mask"], c["name"]
        assert int(to_move[i]) == c["to_move"] and int(status[i]) == c["v_code"]
        assert int(win[i]) == c["win"]
        assert (bytes([1, 2]) + canon[i].astype("<u2").tobytes()).hex() == c["canonical_key"]
        assert int(orbit[i]) == c["orbit"]
        # (expected_player / winner in tests/fixtures.py are test-design metadata; the
        #  values asserted here are what the reference FUNCTIONS returned for the case)
        # (the fixtures' `winner` column is test-design metadata, not check_game_winner's
        #  count-parity rule -- game_utils.py:166-169 -- so it is not asserted here)


def test_dataset_sample1000_keys():
    d = _npz("dataset_depth4_sample1000.npz")
    m = _meta()
    canon = O.canonical(d["states"])
    assert np.array_equal(canon, d["canon192"])
    keys = [bytes([1, 2]) + canon[i].astype("<u2").tobytes() for i in range(len(canon))]
    # SURVEY.md 7.2 / BASELINE.md: sha256 of the 1000 18-byte keys, computed with the reference
    assert hashlib.sha256(b"".join(keys)).hexdigest() == \
        "0a2e95cb49b528c54bbf5b11e55b72c7f71b7cdb388205787f0b1ca72684fa7e" == m["sha256_keys18"]
    assert len(set(keys)) == 1000
    orbit = O.orbit_size(d["states"])
    assert np.array_equal(orbit, d["orbit"]) and int(orbit.astype(int).sum()) == 153592
    mask, to_move, _ = O.movegen_masks(d["states"])
    assert np.array_equal(np.array([bin(int(x)).count("1") for x in mask]), d["legal_moves"])
    assert np.array_equal(to_move, d["side_to_move"])


def test_symmetry_images_and_moves():
    d = _npz("symmetry_images.npz")
    for i in range(len(d["states"])):
        t = d["transform"][i]
        img = O.apply_symmetry(d["states"][i], int(t[0]), bool(t[1]), [int(x) for x in t[2:]])
        assert np.array_equal(img, d["image"][i])
        mv = O.apply_symmetry_to_move(*[int(x) for x in d["move_in"][i]], int(t[0]), bool(t[1]),
                                      [int(x) for x in t[2:]])
        assert list(mv) == d["move_out"][i].tolist()
    for d4 in range(8):
        for sq in range(16):
            assert O.lib().qo_permute16(1 << sq, d4) == 1 << int(d["d4_maps"][d4][sq])


def test_known_canonical_vectors():
    # test_core.py:38-47: empty key;  test_api_portability_report.py:125-285
    empty = np.zeros((1, 8), np.uint16)
    assert O.canonical(empty)[0].tolist() == [0] * 8
    s = np.array([[1, 0, 0, 2048, 0, 2, 64, 0], [1, 0, 0, 0, 0, 0, 0, 0]], np.uint16)
    c = O.canonical(s)
    assert (bytes([1, 2]) + c[0].astype("<u2").tobytes()).hex() == "010200000000000108000400200000000000"
    assert (bytes([1, 2]) + c[1].astype("<u2").tobytes()).hex() == "010200000000000000100000000000000000"
    assert O.orbit_size(s).tolist() == [192, 16]
    mask, tm, _ = O.movegen_masks(s)
    assert [hex(int(x)) for x in mask] == ["0xf7bcb300d580f7bc", "0xfffefffefffeeec0"] and tm.tolist() == [0, 1]
    # SURVEY 8c extra vectors generated with the reference (192 mode | 384 mode)
    vec = [((0, 0, 512, 0, 0, 0, 0, 0), (0, 0, 0, 512, 0, 0, 0, 0), 16, (0, 0, 0, 0, 0, 0, 0, 512)),
           ((0, 8, 0, 0, 0, 0, 4, 0), (0, 0, 0, 4096, 0, 0, 256, 0), 96, (0, 0, 0, 256, 0, 0, 4096, 0)),
           ((4096, 0, 0, 1028, 64, 0, 1, 0), (0, 0, 1280, 8, 0, 1, 0, 512), 192, None),
           ((34816, 640, 4, 10, 80, 1, 8192, 20480), (256, 49152, 4112, 8256, 128, 514, 2056, 1), 192, None)]
    for bb, c192, orb, c384 in vec:
        a = np.array([bb], np.uint16)
        assert tuple(O.canonical(a)[0].tolist()) == c192
        assert int(O.orbit_size(a)[0]) == orb
        if c384 is not None:
            assert tuple(O.canonical(a, True)[0].tolist()) == c384
    # api-portability records written by the reference's own _case_report
    for rec in _meta()["api_portability_cases"]:
        a = np.array([rec["bitboards"]], np.uint16)
        mask, tm, _ = O.movegen_masks(a)
        assert f"0x{int(mask[0]):016x}" == rec["legal_action_mask"] and int(tm[0]) == rec["side_to_move"]
        assert (bytes([1, 2]) + O.canonical(a)[0].astype("<u2").tobytes()).hex() == rec["canonical_key"]
        assert int(O.orbit_size(a)[0]) == rec["orbit_size"]
        win = int(O.win_status(a)[0])
        winner = {0: "none", 1: "player0", 2: "player1"}[win]
        terminal = win != 0 or int(mask[0]) == 0
        if win == 0 and int(mask[0]) == 0:
            winner = "player1" if tm[0] == 0 else "player0"
        assert terminal == rec["terminal"] and winner == rec["winner"]


def test_playouts_philox_vs_reference_rules():
    d = _npz("playouts_philox.npz")
    seed = int(d["seed"])
    n = len(d["empty_outcome"])
    r = O.playouts(np.zeros((1, 8), np.uint16), n, seed, per_playout=True)
    assert np.array_equal(r["outcome"][0], d["empty_outcome"])
    assert np.array_equal(r["plies"][0], d["empty_plies"])
    assert int(r["wins_p0"][0]) == int((d["empty_outcome"] > 0).sum())
    R = d["roots_outcome"].shape[1]
    r = O.playouts(d["roots"], R, seed, first_rollout=int(d["roots_first_rollout"]),
                   root_index_base=int(d["roots_index_base"]), per_playout=True)
    assert np.array_equal(r["outcome"], d["roots_outcome"])
    assert np.array_equal(r["plies"], d["roots_plies"])
    for mp in (16, 12, 5, 0):
        want = d[f"cut{mp}_outcome"]
        r = O.playouts(np.zeros((1, 8), np.uint16), len(want), seed, max_plies=mp, per_playout=True)
        assert np.array_equal(r["outcome"][0], want)
        assert np.array_equal(r["plies"][0], d[f"cut{mp}_plies"])


def test_playout_rate_agrees_with_reference_rng_3sigma():
    # north_star: aggregate win rates within binomial 3 sigma of the reference's own RNG
    m = _meta()
    n = 100000
    r = O.playouts(np.zeros((1, 8), np.uint16), n)
    p_ours = int(r["wins_p0"][0]) / n
    for ref in (m["survey_mt19937_100k"], m["mt19937_empty_board"]):
        p_ref = ref["p0_wins"] / ref["n"]
        sigma = (0.25 / n + 0.25 / ref["n"]) ** 0.5
        assert abs(p_ours - p_ref) < 3 * sigma
    assert int(r["draws"][0]) == 0


def test_symmetry_table_vs_reference_and_published():
    # GAME_TREE_ANALYSIS.md:4-11 (depths 1-8); measured with the reference to depth 5 (reference_meta.json)
    published = [[64, 3, 0, 0, 64], [3392, 51, 0, 0, 3392], [167552, 726, 0, 0, 167552],
                 [6776960, 10946, 0, 6912, 6770048], [231883776, 105632, 1050624, 0, 230833152]]
    assert _meta()["symmetry_table_measured"] == published
    stats, st, mu = O.symmetry_table(5, emit_depth=4, cap=20000)
    assert stats.tolist() == published
    d = _npz("depth4_canonical.npz")
    order = np.lexsort(st.astype("<u2").view(np.uint8).reshape(len(st), 16).T[::-1])
    assert np.array_equal(st[order], d["states"]) and np.array_equal(mu[order], d["mult"])
    assert int(mu.sum()) == 6770048 and len(st) == 10946


def test_perft_raw_matches_table():
    r = O.perft(np.zeros((1, 8), np.uint16), 4)
    assert r["nodes"].tolist() == [1, 64, 3392, 167552, 6776960]
    assert r["wins_p1"].tolist() == [0, 0, 0, 0, 6912] and r["wins_p0"].tolist() == [0] * 5
    # weighted canonical depth-4 roots reproduce depth 5 of the table
    d = _npz("depth4_canonical.npz")
    r = O.perft(d["states"], 1, d["mult"])
    assert int(r["nodes"][0]) == 6770048 and int(r["nodes"][1]) == 231883776 and int(r["wins_p0"][1]) == 1050624


def test_opening_book_summary_vs_reference():
    # tests/golden/opening_book_summary_d5.json was written by the reference's own generator + build_summary
    want = json.load(open(os.path.join(G, "opening_book_summary_d5.json")))
    got = O.opening_book_summary(4)
    assert got.tolist() == [[r["positions"], r["terminal"], r["edges"]] for r in want["per_depth"][:5]]


def test_evaluation_features_and_guided_rollouts_vs_reference():
    # tests/golden/guided_rollouts.npz: evaluation.features / evaluate and playouts under
    # MCTSEngine._select_rollout_move's eval-guided policy, produced with the reference (oracle/gen_golden.py --guided)
    g = _npz("guided_rollouts.npz")
    st = g["states"]
    n = len(st)
    f0, f1 = O.features(st, np.zeros(n, np.uint8)), O.features(st, np.ones(n, np.uint8))
    assert np.array_equal(f0, g["features_p0"]) and np.array_equal(f1, g["features_p1"])
    # floating point: numpy's dot and a left-to-right sum differ by rounding only (tolerance 1e-12 absolute)
    assert np.abs(f0 @ g["fitted_weights"] - g["evaluate_p0_fitted"]).max() < 1e-12
    seed = O.SEED_DEFAULT
    e = np.zeros((1, 8), np.uint16)
    for name, w, eps in (("seeded_eps02", g["seeded_weights"], 0.2), ("fitted_eps02", g["fitted_weights"], 0.2),
                         ("fitted_eps0", g["fitted_weights"], 0.0), ("seeded_eps1", g["seeded_weights"], 1.0)):
        r = O.guided_playouts(e, len(g[name + "_outcome"]), seed, eps, w, root_index_base=3, per_playout=True)
        assert np.array_equal(r["outcome"][0], g[name + "_outcome"]) and np.array_equal(r["plies"][0], g[name + "_plies"]), name
    r = O.guided_playouts(g["roots"], 24, seed, 0.2, g["fitted_weights"], first_rollout=500, root_index_base=40, per_playout=True)
    assert np.array_equal(r["outcome"], g["roots_outcome"]) and np.array_equal(r["plies"], g["roots_plies"])
    r = O.guided_playouts(e, 80, seed, 0.2, g["fitted_weights"], root_index_base=3, max_plies=9, per_playout=True)
    assert np.array_equal(r["outcome"][0], g["cut9_outcome"]) and np.array_equal(r["plies"][0], g["cut9_plies"])


def test_exact_solver_vs_reference_minimax():
    # tests/golden/solver.npz: MinimaxEngine.solve(state).score of the reference on 180 late-game positions
    g = _npz("solver.npz")
    oc, pl, act = O.solve(g["states"])
    assert np.array_equal(oc.astype(float) * (float(g["win"]) - pl), g["score"])
    # the reference's best move is optimal too (its tie-break is its move ordering, ours is the lowest action)
    kids = O.apply_moves(g["states"], g["best_action"])
    koc, kpl, _ = O.solve(kids)
    assert np.array_equal(-koc, oc) and np.array_equal(kpl + 1, pl)
