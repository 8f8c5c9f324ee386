"""GPU tier: the CUDA path, called through the C-ABI (quantik_b200.batch -> libquantik_b200.so),
against (1) golden vectors produced by the reference itself, (2) the pinned C oracle on seeded
inputs, (3) the reference's published tables, (4) size-independent properties at full size.
Everything here is integer work: the bar is bit-exact."""
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle as O
import quantik_b200 as Q
from quantik_b200 import batch as B

G = os.path.join(os.path.dirname(__file__), "golden")
SEED = Q.SEED_DEFAULT


def _npz(name):
    return np.load(os.path.join(G, name))


def _meta():
    return json.load(open(os.path.join(G, "reference_meta.json")))


def _enc(t):
    t = t.astype(np.uint32)
    return (t[:, 0] | (t[:, 1] << 3) | (t[:, 2] << 4) | (t[:, 3] << 6) | (t[:, 4] << 8) | (t[:, 5] << 10)).astype(np.uint16)


@pytest.fixture
def options():
    """qk_set_option switches (kernel variants that compute the same result), restored to the defaults afterwards."""
    from quantik_b200 import _native as N
    touched = set()

    def set_option(name, value):
        touched.add(name)
        N.set_option(name, value)

    yield set_option
    for name in touched:
        N.set_option(name, 0)


def _random_positions(n, seed=1234):
    """reachable positions of every depth + corrupted ones, generated with the oracle"""
    a = O.make_roots(n // 2, seed=seed, min_plies=0, span=13)
    b = a[: n // 4].copy()
    rng = np.random.default_rng(seed)
    b[np.arange(len(b)), rng.integers(0, 8, len(b))] ^= (1 << rng.integers(0, 16, len(b))).astype(np.uint16)
    c = (rng.integers(0, 65536, (n // 4, 8)) & rng.integers(0, 65536, (n // 4, 8)) & rng.integers(0, 65536, (n // 4, 8))).astype(np.uint16)
    return np.ascontiguousarray(np.concatenate([a, b, c]))


# ----------------------------------------------------------------------------- primitives
def test_primitives_vs_reference_goldens():
    d = _npz("states_reference.npz")
    s = d["states"]
    mask, to_move, status = B.movegen_masks(s)
    assert np.array_equal(mask, d["mask"]) and np.array_equal(to_move, d["to_move"]) and np.array_equal(status, d["v_code"])
    assert np.array_equal(B.win_status(s), d["win"])
    c, t = B.canonical_payloads(s, want_transform=True)
    assert np.array_equal(c, d["canon192"]) and np.array_equal(t, _enc(d["tr192"]))
    assert np.array_equal(B.canonical_payloads(s), d["canon192"])
    c, t = B.canonical_payloads(s, color_swap=True, want_transform=True)
    assert np.array_equal(c, d["canon384"]) and np.array_equal(t, _enc(d["tr384"]))
    assert np.array_equal(B.canonical_payloads(s, color_swap=True), d["canon384"])
    assert np.array_equal(B.orbit_sizes(s), d["orbit"])


def test_primitives_vs_oracle_large_random():
    s = _random_positions(200_000)
    mask, to_move, status = B.movegen_masks(s)
    om, ot, os_ = O.movegen_masks(s)
    assert np.array_equal(mask, om) and np.array_equal(to_move, ot) and np.array_equal(status, os_)
    assert len(np.unique(status)) >= 4          # several ValidationResult codes are exercised
    assert np.array_equal(B.win_status(s), O.win_status(s))
    sub = s[:60_000]
    assert np.array_equal(B.canonical_payloads(sub), O.canonical(sub))
    assert np.array_equal(B.canonical_payloads(sub[:20_000], color_swap=True), O.canonical(sub[:20_000], True))
    assert np.array_equal(B.orbit_sizes(sub[:20_000]), O.orbit_size(sub[:20_000]))
    # apply_moves: every legal move of 2000 positions, side to move and explicit player
    valid = s[(status == 0) & (mask != 0)][:2000]
    vm, _, _ = O.movegen_masks(valid)
    rng = np.random.default_rng(5)
    acts = np.array([rng.choice([b for b in range(64) if (int(m) >> b) & 1]) for m in vm], np.uint8)
    assert np.array_equal(B.apply_moves(valid, acts), O.apply_moves(valid, acts))
    explicit = (acts | 0x80 | (rng.integers(0, 2, len(acts)) << 6)).astype(np.uint8)
    assert np.array_equal(B.apply_moves(valid, explicit), O.apply_moves(valid, explicit))


def test_edge_sizes():
    e = np.zeros((0, 8), np.uint16)
    assert len(B.win_status(e)) == 0 and len(B.canonical_payloads(e)) == 0
    one = np.zeros((1, 8), np.uint16)
    assert int(B.movegen_masks(one)[0][0]) == 0xFFFFFFFFFFFFFFFF      # test_api_portability_report.py: empty board
    for n in (1, 31, 32, 33, 255, 257, 1000):                          # ragged tails around warp / block sizes
        s = _random_positions(4 * n)[:n]
        assert np.array_equal(B.canonical_payloads(s), O.canonical(s))
        assert np.array_equal(B.movegen_masks(s)[0], O.movegen_masks(s)[0])


def test_fixture_cases_and_api_records():
    cases = json.load(open(os.path.join(G, "fixture_cases.json")))["cases"]
    s = np.array([c["bb"] for c in cases], np.uint16)
    mask, to_move, status = B.movegen_masks(s)
    win, canon, orbit = B.win_status(s), B.canonical_payloads(s), B.orbit_sizes(s)
    c384 = B.canonical_payloads(s, color_swap=True)
    for i, c in enumerate(cases):
        assert f"0x{int(mask[i]):016x}" == c["mask"] and int(to_move[i]) == c["to_move"] and int(status[i]) == c["v_code"]
        assert int(win[i]) == c["win"] and int(orbit[i]) == c["orbit"]
        assert (b"\x01\x02" + canon[i].astype("<u2").tobytes()).hex() == c["canonical_key"]
        assert c384[i].tolist() == c["canon384"]
    # the reference's cross-stack conformance record (api_portability_report.py:152-191), field by field
    moves = {"empty": (0, 0), "mixed": (2, 12), "single": (1, 5), "single_illegal": (1, 0), "blocked": (0, 1)}
    for rec in _meta()["api_portability_cases"]:
        ours = Q.portability_case_report(rec["case_id"], rec["qfen"], *moves[rec["case_id"]])
        assert ours == rec, rec["case_id"]
    # ... and the whole api-portability-report.v1 document (build_report, :194-226) from one batched pass over the fixture
    fixture = [{"case_id": r["case_id"], "qfen": r["qfen"], "move": {"shape": moves[r["case_id"]][0], "position": moves[r["case_id"]][1]}}
               for r in reversed(_meta()["api_portability_cases"])]
    doc = Q.portability_report(fixture, contracts_release="test", contract_ids={"qfen": "qfen.v1"})
    assert doc["schema"] == "api-portability-report.v1" and doc["contracts_release"] == "test" and doc["contract_ids"] == {"qfen": "qfen.v1"}
    assert doc["cases"] == sorted(_meta()["api_portability_cases"], key=lambda r: r["case_id"])
    json.dumps(doc, sort_keys=True)
    for bad in ([], [{"case_id": "", "qfen": "..../..../..../....", "move": {"shape": 0, "position": 0}}],
                [{"case_id": "x", "qfen": "..../..../..../....", "move": {"shape": 4, "position": 0}}],
                [{"case_id": "x", "qfen": "AA../..../..../....", "move": {"shape": 0, "position": 5}}]):      # invalid game state
        with pytest.raises(ValueError):
            Q.portability_report(bad)


def test_dataset_sample1000_sha256_and_orbit_dedup():
    d = _npz("dataset_depth4_sample1000.npz")
    keys = B.canonical_keys(d["states"])
    assert hashlib.sha256(b"".join(keys)).hexdigest() == "0a2e95cb49b528c54bbf5b11e55b72c7f71b7cdb388205787f0b1ca72684fa7e"
    assert len(set(keys)) == 1000
    assert int(np.asarray(B.orbit_sizes(d["states"])).astype(int).sum()) == 153592
    # the full 192-orbit of every position dedups back to the same 1000 payloads, 192 each
    codes = np.array([B.encode_transform(d4, False, O.perm(pi)) for d4 in range(8) for pi in range(24)], np.uint16)
    images = B.apply_symmetry(np.repeat(d["states"], 192, axis=0), np.tile(codes, 1000))
    assert len(np.unique(images.view(np.uint8).reshape(-1, 16), axis=0)) == 153592
    uniq, mult = B.canon_dedup(images)
    want = d["canon192"][np.lexsort(d["canon192"].astype("<u2").view(np.uint8).reshape(-1, 16).T[::-1])]
    assert np.array_equal(uniq, want) and np.all(mult == 192)


def test_symmetry_images_and_single_state_api():
    g = _npz("symmetry_images.npz")
    assert np.array_equal(B.apply_symmetry(g["states"], _enc(g["transform"])), g["image"])
    # single-state mirror of the reference API
    st = Q.State.from_qfen("A.bC/..../d..B/...a")
    assert st.canonical_key().hex() == "010200000001010008000008020000000010" and st.symmetry_count() == 192
    assert Q.State.empty().canonical_key() == b"\x01\x02" + b"\x00" * 16           # tests/test_core.py:38-47
    bb = (1, 0, 0, 2048, 0, 2, 64, 0)
    player, by_shape = Q.generate_legal_moves(bb)
    assert player == 0 and sorted(by_shape) == [0, 1, 2, 3] and sum(len(v) for v in by_shape.values()) == 35
    assert Q.generate_legal_moves(bb, player_id=1) == (0, {})                        # move.py:228-229
    assert Q.generate_legal_moves((1, 1, 0, 0, 0, 0, 0, 0)) == (0, {})               # move.py:223-225
    assert Q.apply_move(bb, Q.Move(0, 2, 12)) == (1, 0, 4096, 2048, 0, 2, 64, 0)
    assert Q.check_game_winner(Q.bb_from_qfen("AbCd/..../..../....")) == Q.WinStatus.PLAYER_1_WINS
    assert Q.has_winning_line(bb) is False
    canon, tr = Q.SymmetryHandler.find_canonical_form((0, 8, 0, 0, 0, 0, 4, 0))
    assert canon == (0, 0, 0, 4096, 0, 0, 256, 0) and (tr.d4_index, tr.color_swap, tr.shape_perm) == (6, False, (0, 3, 2, 1))
    assert Q.SymmetryHandler.apply_symmetry((0, 8, 0, 0, 0, 0, 4, 0), tr) == canon
    # forbidden squares for an opponent piece on square 0 (tests/test_move.py:225-276)
    _, moves = Q.generate_legal_moves((1, 0, 0, 0, 0, 0, 0, 0))
    assert sorted(set(range(16)) - {m.position for m in moves[0]}) == [0, 1, 2, 3, 4, 5, 8, 12]
    # first moves: 64 -> 3 canonical classes of multiplicity 16/16/32 (tests/test_game_stats.py:40-73)
    firsts = np.array([Q.apply_move((0,) * 8, Q.Move(0, s, p)) for s in range(4) for p in range(16)], np.uint16)
    uniq, mult = B.canon_dedup(firsts)
    assert len(uniq) == 3 and sorted(mult.tolist()) == [16, 16, 32]


def test_apply_symmetry_to_moves_batched():
    """qk_apply_symmetry_to_moves against the reference's own images (apply_symmetry_to_move, symmetry.py:425-464), the
    oracle on every (transform, move) pair, and the property that makes it useful: moving then mapping = mapping then moving."""
    import itertools
    g = _npz("symmetry_images.npz")
    mi, mo = g["move_in"].astype(np.int64), g["move_out"].astype(np.int64)
    pl, ac = B.apply_symmetry_to_moves(mi[:, 0], mi[:, 1] * 16 + mi[:, 2], _enc(g["transform"]))
    assert np.array_equal(pl, mo[:, 0]) and np.array_equal(ac, mo[:, 1] * 16 + mo[:, 2])
    # all 384 transforms x 128 moves against the oracle's scalar restatement, and the inverse map undoes every one
    codes, players, actions, want = [], [], [], []
    for d4, cs, perm in itertools.product(range(8), (False, True), itertools.permutations(range(4))):
        for player, a in itertools.product((0, 1), range(64)):
            codes.append(B.encode_transform(d4, cs, perm)); players.append(player); actions.append(a)
            want.append(O.apply_symmetry_to_move(player, a >> 4, a & 15, d4, cs, perm))
    codes, players, actions, want = np.array(codes, np.uint16), np.array(players, np.uint8), np.array(actions, np.uint8), np.array(want)
    pl, ac = B.apply_symmetry_to_moves(players, actions, codes)
    assert np.array_equal(pl, want[:, 0]) and np.array_equal(ac, want[:, 1] * 16 + want[:, 2])
    bp, ba = B.apply_symmetry_to_moves(pl, ac, codes, inverse=True)
    assert np.array_equal(bp, players) and np.array_equal(ba, actions)
    # book lookup round trip: canonicalise a position, pick a legal move THERE, map it back with inverse=True: it is legal
    # on the original position and leads to the same canonical class
    roots = O.make_roots(5000, seed=21, min_plies=0, span=12)
    canon, tr = B.canonical_payloads(roots, want_transform=True)
    cmask, cmover, _ = B.movegen_masks(canon)
    keep = cmask != 0
    roots, canon, tr, cmask, cmover = roots[keep], canon[keep], tr[keep], cmask[keep], cmover[keep]
    first = np.array([(int(m) & -int(m)).bit_length() - 1 for m in cmask], np.uint8)           # lowest legal action on the canonical board
    bp, ba = B.apply_symmetry_to_moves(cmover, first, tr, inverse=True)
    rmask, rmover, _ = B.movegen_masks(roots)
    assert np.array_equal(bp, rmover) and np.all((rmask >> ba.astype(np.uint64)) & np.uint64(1))
    assert np.array_equal(B.canonical_payloads(B.apply_moves(roots, ba)), B.canonical_payloads(B.apply_moves(canon, first)))


def test_canonical_properties_at_full_size():
    """size-independent properties on 2^23 device-resident states (tests/test_core.py:159-181 at scale):
    idempotence, invariance under every symmetry, orbit-size invariance, multiplicity conservation"""
    import torch
    n = 1 << 23
    base = B.random_roots(1 << 18, seed=77, min_plies=0, span=14, as_torch=True)
    states = base.repeat(n // (1 << 18), 1)
    g = torch.Generator(device="cuda").manual_seed(5)
    d4 = torch.randint(0, 8, (n,), device="cuda", generator=g)
    perms = torch.tensor([B.encode_transform(0, False, O.perm(i)) for i in range(24)], device="cuda")
    codes = (perms[torch.randint(0, 24, (n,), device="cuda", generator=g)] | d4).to(torch.int16)
    images = B.apply_symmetry(states, codes)
    c0, c1 = B.canonical_payloads(states), B.canonical_payloads(images)
    assert torch.equal(c0, c1)                                   # invariance: one key per orbit
    assert torch.equal(B.canonical_payloads(c0), c0)             # idempotence
    assert torch.equal(B.orbit_sizes(states), B.orbit_sizes(images))
    # 384-image mode: additionally invariant under the colour swap
    swapped = B.apply_symmetry(states[: 1 << 20], (codes[: 1 << 20] | 8))
    assert torch.equal(B.canonical_payloads(states[: 1 << 20], color_swap=True), B.canonical_payloads(swapped, color_swap=True))
    # dedup conserves multiplicity and returns exactly the distinct canonical payloads
    uniq, mult = B.canon_dedup(images, cap=1 << 19)
    assert int(mult.sum()) == n
    ref = torch.unique(c0.view(torch.int64).reshape(-1, 2), dim=0)
    assert len(uniq) == len(ref)
    # spot-check against the oracle
    idx = torch.randint(0, n, (4096,), device="cuda", generator=g).cpu().numpy()
    assert np.array_equal(c1.cpu().numpy().view(np.uint16)[idx], O.canonical(images.cpu().numpy().view(np.uint16)[idx]))


def test_canon_dedup_vs_oracle_with_multiplicities():
    s = _random_positions(120_000, seed=77)[:60_000]      # reachable states: many symmetric duplicates
    mult = np.random.default_rng(3).integers(1, 1000, len(s)).astype(np.uint64)
    uniq, um = B.canon_dedup(s, mult)
    oc = O.canonical(s).astype("<u2").view(np.uint8).reshape(-1, 16)
    keys, inv = np.unique(oc, axis=0, return_inverse=True)
    want_mult = np.zeros(len(keys), np.uint64)
    np.add.at(want_mult, inv.ravel(), mult)
    assert np.array_equal(uniq.astype("<u2").view(np.uint8).reshape(-1, 16), keys)
    assert np.array_equal(um, want_mult)
    u2, m2 = B.canon_dedup(s, mult, sort_output=False)      # QK_DEDUP_UNSORTED: same set, table order
    order = np.lexsort(u2.astype("<u2").view(np.uint8).reshape(-1, 16).T[::-1])
    assert np.array_equal(u2[order], uniq) and np.array_equal(m2[order], um)
    for n_small in (1, 2, 1023, 1025, 4097):               # sort sizes around the shared-memory tile
        u3, m3 = B.canon_dedup(s[:n_small], mult[:n_small])
        k3, inv3 = np.unique(oc[:n_small], axis=0, return_inverse=True)
        w3 = np.zeros(len(k3), np.uint64)
        np.add.at(w3, inv3.ravel(), mult[:n_small])
        assert np.array_equal(u3.astype("<u2").view(np.uint8).reshape(-1, 16), k3) and np.array_equal(m3, w3)
    ones = np.full((5, 8), 0xFFFF, np.uint16)              # the payload that equals the table's empty marker
    uniq, um = B.canon_dedup(np.concatenate([s[:100], ones]))
    assert np.array_equal(uniq[-1], ones[0]) and int(um[-1]) == 5


def test_two_level_merge_equals_one_level_table(options):
    """qk_canon_dedup's two paths -- one table (the reference's one dictionary, game_stats.py:487-505) and keys routed into
    L2-sized parts that are merged one at a time -- give the same classes and multiplicities: small inputs (forced), weighted,
    colour-swap mode, the all-ones payload, sorted and table order, and a 2^22-key set where the automatic choice is two-level."""
    import torch
    rng = np.random.default_rng(11)
    small = _random_positions(30_000, seed=3)
    small = np.concatenate([small, small[:7000], np.full((3, 8), 0xFFFF, np.uint16)])
    wts = rng.integers(1, 1 << 40, len(small)).astype(np.uint64)
    big = B.random_roots(1 << 22, seed=99, min_plies=2, span=8)
    for states, mult, cs in ((small, None, False), (small, wts, False), (small, wts, True), (big, None, False)):
        res = {}
        for mode in (1, 2):
            options("dedup_mode", mode)
            u, m = B.canon_dedup(states, mult, color_swap=cs)
            res[mode] = (u, m)
            tu, tm = B.canon_dedup(torch.from_numpy(states.view(np.int16)).cuda(), None if mult is None else torch.from_numpy(mult.view(np.int64)).cuda(),
                                   color_swap=cs, sort_output=False)
            tu, tm = tu.cpu().numpy().view(np.uint16), tm.cpu().numpy().view(np.uint64)
            order = np.lexsort(tu.astype("<u2").view(np.uint8).reshape(-1, 16).T[::-1])
            assert np.array_equal(tu[order], u) and np.array_equal(tm[order], m), (mode, "unsorted path")
        assert np.array_equal(res[1][0], res[2][0]) and np.array_equal(res[1][1], res[2][1])
        assert int(res[2][1].sum()) == (len(states) if mult is None else int(mult.sum()))
    options("dedup_mode", 0)
    u0, m0 = B.canon_dedup(big)
    assert np.array_equal(u0, res[1][0]) and np.array_equal(m0, res[1][1])


# ----------------------------------------------------------------------------- playouts
def test_playouts_vs_reference_rules_goldens():
    p = _npz("playouts_philox.npz")
    seed = int(p["seed"])
    e = np.zeros((1, 8), np.uint16)
    n = len(p["empty_outcome"])
    r = B.playouts(e, n, seed, per_playout=True)
    assert np.array_equal(r["outcome"][0], p["empty_outcome"]) and np.array_equal(r["plies"][0], p["empty_plies"])
    assert int(r["wins_p0"][0]) == int((p["empty_outcome"] > 0).sum()) and int(r["wins_p1"][0]) == int((p["empty_outcome"] < 0).sum())
    r = B.playouts(p["roots"], p["roots_outcome"].shape[1], seed, first_rollout=int(p["roots_first_rollout"]),
                   root_index_base=int(p["roots_index_base"]), per_playout=True)
    assert np.array_equal(r["outcome"], p["roots_outcome"]) and np.array_equal(r["plies"], p["roots_plies"])
    assert np.array_equal(r["wins_p0"], (p["roots_outcome"] > 0).sum(1))
    for mp in (16, 12, 5, 0):                                # MCTS cut-off semantics (mcts.py:412,436-437)
        want = p[f"cut{mp}_outcome"]
        r = B.playouts(e, len(want), seed, max_plies=mp, per_playout=True)
        assert np.array_equal(r["outcome"][0], want) and np.array_equal(r["plies"][0], p[f"cut{mp}_plies"])
        assert int(r["draws"][0]) == int((want == 0).sum())


def test_playouts_vs_oracle_per_seed_1m():
    n = 1 << 20
    e = np.zeros((1, 8), np.uint16)
    g = B.playouts(e, n, SEED, per_playout=True)
    o = O.playouts(e, n, SEED, per_playout=True)
    assert np.array_equal(g["outcome"], o["outcome"]) and np.array_equal(g["plies"], o["plies"])
    assert int(g["wins_p0"][0]) == int(o["wins_p0"][0]) and int(g["wins_p1"][0]) == int(o["wins_p1"][0]) and int(g["draws"][0]) == 0
    # tallies-only kernel variant gives the same counts
    t = B.playouts(e, n, SEED)
    assert int(t["wins_p0"][0]) == int(o["wins_p0"][0]) and int(t["wins_p1"][0]) == int(o["wins_p1"][0])
    # north_star: win rate within binomial 3 sigma of the reference's own RNG (MT19937 random.choice)
    for ref in (_meta()["survey_mt19937_100k"], _meta()["mt19937_empty_board"]):
        sigma = (0.25 / n + 0.25 / ref["n"]) ** 0.5
        assert abs(int(t["wins_p0"][0]) / n - ref["p0_wins"] / ref["n"]) < 3 * sigma
    lens = np.bincount(g["plies"][0], minlength=17)
    assert lens[:4].sum() == 0 and lens[16] > 0             # games last 4..16 plies


def test_playouts_sharding_invariance_and_mid_roots():
    roots = O.make_roots(300, seed=SEED)
    assert np.array_equal(B.random_roots(300, seed=SEED), roots)
    assert np.array_equal(B.random_roots(4096, seed=99, first=1000), O.make_roots(4096, seed=99, first=1000))
    full = B.playouts(roots, 256, SEED, per_playout=True)
    o = O.playouts(roots, 256, SEED, per_playout=True)
    assert np.array_equal(full["outcome"], o["outcome"]) and np.array_equal(full["plies"], o["plies"])
    assert np.array_equal(full["wins_p0"], o["wins_p0"]) and np.array_equal(full["wins_p1"], o["wins_p1"])
    # a root shard / rollout shard computed separately is identical to the slice of the full job
    part = B.playouts(roots[100:200], 256, SEED, root_index_base=100, per_playout=True)
    assert np.array_equal(part["outcome"], full["outcome"][100:200])
    half = B.playouts(roots, 128, SEED, first_rollout=128, per_playout=True)
    assert np.array_equal(half["outcome"], full["outcome"][:, 128:])
    # special roots: blocked mover, already won, invalid
    sp = np.array([Q.bb_from_qfen("A..C/bbd./CD.A/.adB"), Q.bb_from_qfen("AbCd/..../..../...."), (1, 1, 0, 0, 0, 0, 0, 0)], np.uint16)
    g, oo = B.playouts(sp, 64, SEED, per_playout=True), O.playouts(sp, 64, SEED, per_playout=True)
    assert np.array_equal(g["outcome"], oo["outcome"]) and np.array_equal(g["plies"], oo["plies"])
    assert g["wins_p0"].tolist() == [64, 0, 0] and g["wins_p1"].tolist() == [0, 64, 64]
    # cut-off on mid-game roots
    g, oo = B.playouts(roots[:64], 128, SEED, max_plies=3, per_playout=True), O.playouts(roots[:64], 128, SEED, max_plies=3, per_playout=True)
    assert np.array_equal(g["outcome"], oo["outcome"]) and np.array_equal(g["draws"], oo["draws"])


def test_playouts_config4_sample_and_full_size_properties():
    roots = B.random_roots(65536, seed=SEED)
    assert np.array_equal(roots, O.make_roots(65536, seed=SEED))
    mask, _, status = B.movegen_masks(roots)
    assert np.all(status == 0) and np.all(mask != 0) and np.all(B.win_status(roots) == 0)
    pieces = np.unpackbits(np.ascontiguousarray(roots[:4096]).view(np.uint8), axis=1).sum(1)
    assert pieces.min() == 5 and pieces.max() == 11
    r = B.playouts(roots, 1024, SEED)                          # BASELINE config 4: 67 108 864 playouts
    assert np.all(r["wins_p0"].astype(np.int64) + r["wins_p1"] == 1024) and np.all(r["draws"] == 0)
    # "bit-exact vs C++ oracle on ALL roots" -- every one of the 65 536 tallies (about a minute of
    # oracle time on 16 host threads), in slices so a mismatch names its root
    O.set_num_threads(len(os.sched_getaffinity(0)))
    for b in range(0, 65536, 8192):
        o = O.playouts(roots[b:b + 8192], 1024, SEED, root_index_base=b)
        for k in ("wins_p0", "wins_p1", "draws"):
            bad = np.flatnonzero(np.asarray(r[k][b:b + 8192]) != o[k])
            assert len(bad) == 0, (k, b + int(bad[0]))
    # BASELINE config 1 scaled up: 2^26 playouts from the empty board
    n = 1 << 26
    t = B.playouts(np.zeros((1, 8), np.uint16), n, SEED)
    assert int(t["wins_p0"][0]) + int(t["wins_p1"][0]) == n and int(t["draws"][0]) == 0
    assert abs(int(t["wins_p0"][0]) / n - 0.4867) < 0.004      # reference MT19937 estimate 0.4853..0.4881


# ----------------------------------------------------------------------------- perft
PUBLISHED = [  # GAME_TREE_ANALYSIS.md:4-9: total legal moves, P0 wins, P1 wins, ongoing
    [64, 0, 0, 64], [3392, 0, 0, 3392], [167552, 0, 0, 167552], [6776960, 0, 6912, 6770048],
    [231883776, 1050624, 0, 230833152], [6241600512, 0, 81653760, 6159946752]]


def test_perft_empty_board_depth6_published_table():
    e = np.zeros((1, 8), np.uint16)
    for depth in range(0, 7):
        p = B.perft(e, depth)
        assert int(p["nodes"][0]) == 1
        assert B.symmetry_table_rows(p) == PUBLISHED[:depth], depth
        assert int(p["blocked_p0_loses"].sum()) == 0 and int(p["blocked_p1_loses"].sum()) == 0


def test_perft_weighted_canonical_roots_and_oracle():
    d = _npz("depth4_canonical.npz")                         # the 10 946 depth-4 classes from the reference
    p = B.perft(d["states"], 2, d["mult"])
    assert int(p["nodes"][0]) == 6770048
    assert [int(p["nodes"][1]), int(p["wins_p0"][1]), int(p["wins_p1"][1])] == [231883776, 1050624, 0]
    assert [int(p["nodes"][2]), int(p["wins_p0"][2]), int(p["wins_p1"][2])] == [6241600512, 0, 81653760]
    # depth 7 and 8 of the published table through symmetry-weighted roots (GAME_TREE_ANALYSIS.md:10-11)
    p = B.perft(d["states"], 4, d["mult"])
    assert int(p["nodes"][3]) == 132400548864 and int(p["wins_p0"][3]) + int(p["blocked_p1_loses"][2]) == 3886838784
    assert int(p["nodes"][4]) == 2097048766464 and int(p["wins_p1"][4]) + int(p["blocked_p0_loses"][3]) == 118862401536
    assert int(p["nodes"][4]) - int(p["wins_p0"][4]) - int(p["wins_p1"][4]) == 1978186364928
    # late-game roots exercise the blocked-mover counters; compare all five arrays with the oracle
    roots = O.make_roots(64, seed=4242, min_plies=8, span=4)
    mult = np.arange(1, 65, dtype=np.uint64)
    for depth in (1, 2, 3, 5):
        g, o = B.perft(roots, depth, mult), O.perft(roots, depth, mult)
        for k in o:
            assert np.array_equal(g[k], o[k]), (depth, k)
    assert int(O.perft(roots, 5, mult)["blocked_p0_loses"].sum() + O.perft(roots, 5, mult)["blocked_p1_loses"].sum()) > 0
    # won and invalid roots
    sp = np.array([Q.bb_from_qfen("AbCd/..../..../...."), (1, 1, 0, 0, 0, 0, 0, 0), Q.bb_from_qfen("A..C/bbd./CD.A/.adB")], np.uint16)
    g, o = B.perft(sp, 3), O.perft(sp, 3)
    for k in o:
        assert np.array_equal(g[k], o[k]), k


def test_perft_lane_kernel_matches_warp_kernel_and_oracle(options):
    e = np.zeros((1, 8), np.uint16)
    roots = O.make_roots(64, seed=4242, min_plies=8, span=4)
    mult = np.arange(1, 65, dtype=np.uint64)
    mid = O.make_roots(3000, seed=17, min_plies=4, span=5)
    ref6 = None
    for mode in ("tile", "tile-generic", "tile-dealt-only", "tile-own-lists", "tile-own-longest", "tile-own-mean", "warp", "lane", "tile-small"):
        options("perft_kernel", {"tile": 1, "lane": 2, "warp": 3}[mode[:4]])
        # leaf parents: everything through the dealt stage / lanes evaluate children of their own as soon as one lane
        # of the warp gives, as many as the shortest / longest / mean list of the warp (default: wide trees only)
        options("perft_own_lanes", {"tile-dealt-only": 33, "tile-own-lists": 1, "tile-own-longest": 65, "tile-own-mean": 129}.get(mode, 0))
        # tile-generic: 64-bit multiplicities / mixed sides-to-move form, on inputs that allow the fast one
        options("perft_tile_generic", 1 if mode == "tile-generic" else 0)
        if mode == "tile-small":        # the depth-first kernel straight on the 64 mixed-parity roots: every frame depth
            options("perft_min_items", 1)
        p = B.perft(e, 6)
        assert B.symmetry_table_rows(p) == PUBLISHED, mode
        for depth in (2, 3, 5, 7):
            g, o = B.perft(roots, depth, mult), O.perft(roots, depth, mult)
            for k in o:
                assert np.array_equal(g[k], o[k]), (mode, depth, k)
        g = B.perft(mid, 4)
        if mode == "tile-small":        # same-side roots with multiplicities beyond 32 bits: generic form, exact sums
            same = roots[O.validate(roots)[0] == 0][:8]
            big = (np.arange(1, len(same) + 1, dtype=np.uint64) << np.uint64(33)) + np.uint64(7)
            gb, ob = B.perft(same, 3, big), O.perft(same, 3, big)
            for k in ob:
                assert np.array_equal(gb[k], ob[k]), (mode, "big multiplicities", k)
        ref6 = ref6 or g
        for k in g:
            assert np.array_equal(g[k], ref6[k]), (mode, k)
    options("perft_kernel", 0)
    options("perft_min_items", 0)
    options("perft_own_lanes", 0)
    # the automatic choice (lane kernel below wide frontiers) on a deep job: depth-4 classes + 6 = depth 10
    d = _npz("depth4_canonical.npz")
    p = B.perft(d["states"], 6, d["mult"])
    rows = B.symmetry_table(10)
    for r in range(1, 7):
        want = rows[4 + r - 1]
        assert [int(p["nodes"][r]), int(p["wins_p0"][r]) + int(p["blocked_p1_loses"][r - 1]),
                int(p["wins_p1"][r]) + int(p["blocked_p0_loses"][r - 1]),
                int(p["nodes"][r]) - int(p["wins_p0"][r]) - int(p["wins_p1"][r])] == [want[0], want[2], want[3], want[4]]


def test_perft_is_symmetry_invariant():
    roots = O.make_roots(16, seed=31, min_plies=4, span=3)
    base = B.perft(roots, 4)
    codes = np.full(16, B.encode_transform(3, False, (2, 0, 3, 1)), np.uint16)
    img = B.perft(B.apply_symmetry(roots, codes), 4)
    for k in base:
        assert np.array_equal(base[k], img[k])


def test_part_streams_and_table_sizes_do_not_change_the_merge(options):
    """The two-level merge with 1-4 part tables in flight (side streams) and part tables from 1 MiB (many parts) to larger
    than the input needs (two parts, fewer than the streams) returns the same classes and multiplicities as the one table."""
    states = B.random_roots(1 << 20, seed=5, min_plies=2, span=7)
    wts = np.random.default_rng(2).integers(1, 1 << 33, len(states)).astype(np.uint64)
    options("dedup_mode", 1)
    want = B.canon_dedup(states, wts)
    options("dedup_mode", 2)
    for streams, mb in ((1, 1), (2, 1), (3, 4), (4, 16), (4, 512), (2, 0)):
        options("dedup_streams", streams)
        options("dedup_part_mb", mb)
        u, m = B.canon_dedup(states, wts)
        assert np.array_equal(u, want[0]) and np.array_equal(m, want[1]), (streams, mb)


def test_small_call_staging_block_is_transparent(options):
    """Host buffers of small calls travel through the per-device staging block (pinned + device, 256 KiB); larger ones
    and calls made with the block switched off go through the memory pool.  Same results either way, also for a call
    whose buffers only partly fit, and for device-buffer calls on two streams that share the block for their scratch."""
    import torch
    small, big = _random_positions(500, seed=21), _random_positions(40_000, seed=22)      # 8 KB / 640 KB of states
    e = np.zeros((1, 8), np.uint16)
    res = {}
    for stage in (1, 0):
        options("call_stage", stage)
        res[stage] = (B.movegen_masks(small), B.canonical_payloads(small), B.win_status(big), B.canonical_payloads(big),
                      B.playouts(e, 4096, seed=SEED), B.perft(e, 4), B.canon_dedup(small))
    for a, b in zip(res[0], res[1]):
        if isinstance(a, dict):
            assert all(np.array_equal(a[k], b[k]) for k in a)
        elif isinstance(a, tuple):
            assert all(np.array_equal(x, y) for x, y in zip(a, b))
        else:
            assert np.array_equal(a, b)
    # two streams, device buffers: the stage's scratch is handed from one stream to the other through an event
    options("call_stage", 0)
    roots = torch.from_numpy(B.random_roots(64, seed=3, min_plies=3, span=6).view(np.int16)).cuda()
    want = B.playouts(roots, 512, seed=SEED)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    outs = []
    for rep in range(6):
        with torch.cuda.stream(s1 if rep % 2 == 0 else s2):
            outs.append(B.playouts(roots, 512, seed=SEED))
    torch.cuda.synchronize()
    for o in outs:
        assert all(np.array_equal(np.asarray(o[k].cpu() if hasattr(o[k], "cpu") else o[k]), np.asarray(want[k].cpu() if hasattr(want[k], "cpu") else want[k])) for k in ("wins_p0", "wins_p1", "draws"))


# ----------------------------------------------------------------------------- canonical BFS
def test_symmetry_table_bfs_reproduces_published_table_with_unique_counts(options):
    # GAME_TREE_ANALYSIS.md:4-11: total legal moves, unique canonical, P0 wins, P1 wins, ongoing
    table = [[64, 3, 0, 0, 64], [3392, 51, 0, 0, 3392], [167552, 726, 0, 0, 167552], [6776960, 10946, 0, 6912, 6770048],
             [231883776, 105632, 1050624, 0, 230833152], [6241600512, 901916, 0, 81653760, 6159946752],
             [132400548864, 4658465, 3886838784, 0, 128513710080],
             [2097048766464, 17900160, 0, 118862401536, 1978186364928]]
    assert B.symmetry_table(8) == table
    assert B.symmetry_table(5, on_device=False) == table[:5]          # host-buffer path of the same call
    for mode in (1, 2):       # every level through the one table / through the routed (two-level) form, whatever its size
        options("dedup_mode", mode)
        assert B.symmetry_table(8) == table, mode
        assert B.symmetry_table(4, on_device=False, sort_output=True) == table[:4], mode
    options("dedup_mode", 0)
    assert [r[1] for r in table] == [3, 51, 726, 10946, 105632, 901916, 4658465, 17900160]   # beam_search.py:210-219
    # the depth-4 frontier equals the reference's own (states and multiplicities)
    f, m = np.zeros((1, 8), np.uint16), np.ones(1, np.uint64)
    for _ in range(4):
        f, m, _ = B.bfs_level(f, m)
    d = _npz("depth4_canonical.npz")
    assert np.array_equal(f, d["states"]) and np.array_equal(m, d["mult"])


@pytest.mark.parametrize("dedup_mode", [0, 2])
def test_bfs_level_vs_oracle_composition(options, dedup_mode):
    options("dedup_mode", dedup_mode)
    parents = np.unique(O.canonical(O.make_roots(3000, seed=5, min_plies=6, span=6)), axis=0)
    mult = np.random.default_rng(8).integers(1, 50, len(parents)).astype(np.uint64)
    f, m, st = B.bfs_level(parents, mult)
    mask, to_move, _ = O.movegen_masks(parents)
    want, tot, w = {}, 0, [0, 0]
    blocked = [0, 0]
    for i in range(len(parents)):
        mk = int(mask[i])
        if mk == 0:
            blocked[int(to_move[i])] += int(mult[i])
            continue
        acts = np.array([b for b in range(64) if (mk >> b) & 1], np.uint8)
        kids = O.apply_moves(np.repeat(parents[i:i + 1], len(acts), 0), acts)
        wins = O.win_status(kids)
        tot += len(acts) * int(mult[i])
        w[int(to_move[i])] += int((wins != 0).sum()) * int(mult[i])
        for c in O.canonical(kids[wins == 0]):
            want[c.astype("<u2").tobytes()] = want.get(c.astype("<u2").tobytes(), 0) + int(mult[i])
    assert st == {"total_moves": tot, "line_wins_p0": w[0], "line_wins_p1": w[1], "blocked_p0": blocked[0], "blocked_p1": blocked[1]}
    keys = sorted(want)
    assert [r.astype("<u2").tobytes() for r in f] == keys and m.tolist() == [want[k] for k in keys]
    f2, m2, _ = B.bfs_level(parents, mult, cap=16)               # forces the grow-and-retry path
    assert np.array_equal(f2, f) and np.array_equal(m2, m)


def test_opening_book_summary_contract():
    # opening-book-summary.v1 written by the reference itself (generate_opening_book.generate -> sqlite ->
    # opening_book_summary.build_summary, depth 5, 181 s) must be reproduced field for field
    want = json.load(open(os.path.join(G, "opening_book_summary_d5.json")))
    assert B.opening_book_summary(5) == want
    assert B.opening_book_summary(3) == {**{k: want[k] for k in ("schema", "contract_version")}, "depth": 3,
                                         "total_positions": 781, "terminal_positions": 0, "total_edges": 22998,
                                         "per_depth": want["per_depth"][:4]}
    # non-terminal classes per depth are exactly SymmetryTable's unique canonical states
    s = B.opening_book_summary(8)
    uniq = [3, 51, 726, 10946, 105632, 901916, 4658465, 17900160]
    rows = B.symmetry_table(8)
    for d in range(1, 9):
        won = s["per_depth"][d]["positions"] - uniq[d - 1]
        assert won >= 0 and s["per_depth"][d]["terminal"] >= won
        assert rows[d - 1][1] == uniq[d - 1]


def test_abi_error_paths_and_degenerate_sizes():
    import ctypes
    from quantik_b200 import _native as N
    lib = N.init(0)
    s = np.zeros((4, 8), np.uint16)
    out = np.zeros(4, np.uint8)
    assert lib.qk_win_status(None, 4, out.ctypes.data, 0, None) == -1 and b"NULL" in lib.qk_last_error()
    assert lib.qk_win_status(s.ctypes.data, 4, out.ctypes.data, 7, None) == -3          # device never initialised
    assert lib.qk_init(99) == -1
    arrs = [np.zeros(20, np.uint64) for _ in range(5)]
    assert lib.qk_perft(s.ctypes.data, None, 4, 17, *[a.ctypes.data for a in arrs], 0, None) == -1   # depth > 16
    assert lib.qk_random_roots(1, 0, 4, 10, 8, s.ctypes.data, 0, None) == -1             # 10 + 8 - 1 > 15 plies
    with pytest.raises(Q.QuantikB200Error):
        B.canon_dedup(O.make_roots(500), cap=3)                                          # more classes than cap
    with pytest.raises(ValueError):
        B.apply_moves(s, np.zeros(3, np.uint8))
    with pytest.raises(ValueError):
        B.movegen_masks(np.zeros(7, np.uint16))
    # the exchange entry points: argument checks, and an empty share still publishes its (zero) counts
    import ctypes
    import torch
    box = torch.full((B.inbox_bytes(2, 64),), 0xFF, dtype=torch.uint8, device="cuda:0")
    ptrs = (ctypes.c_void_p * 2)(box.data_ptr(), box.data_ptr())
    sent = np.full(2, 7, np.uint64)
    assert lib.qk_canon_dedup_scatter(s.ctypes.data, None, 4, 0, ptrs, 17, 0, 64, sent.ctypes.data, 0, None) == -1   # world > 16
    assert lib.qk_canon_dedup_scatter(s.ctypes.data, None, 4, 0, ptrs, 2, 2, 64, sent.ctypes.data, 0, None) == -1    # rank >= world
    assert lib.qk_canon_dedup_scatter(s.ctypes.data, None, 4, 0, None, 2, 0, 64, sent.ctypes.data, 0, None) == -1    # no inboxes
    assert lib.qk_canon_dedup_scatter(s.ctypes.data, None, 4, 0, ptrs, 2, 0, 0, sent.ctypes.data, 0, None) == -1     # empty regions
    nb = ctypes.c_size_t(0)
    assert lib.qk_inbox_bytes(0, 64, ctypes.byref(nb)) == -1 and lib.qk_inbox_bytes(2, 64, None) == -1
    one = (ctypes.c_void_p * 1)(box.data_ptr())
    assert B.canon_dedup_scatter(np.zeros((0, 8), np.uint16), None, [box.data_ptr()], 0, 64).tolist() == [0]
    assert int(box[:8].view(torch.int64).item()) == 0
    u_, m_ = B.inbox_merge(box.data_ptr(), 1, 64, 1)
    assert len(u_) == 0 and len(m_) == 0
    assert B.canon_dedup_scatter(s, None, [box.data_ptr()], 0, 64).tolist() == [1]          # four empty boards: one class
    u_, m_ = B.inbox_merge(box.data_ptr(), 1, 64, 1)
    assert u_.cpu().numpy().view(np.uint16).tolist() == [[0] * 8] and m_.tolist() == [4]
    del one
    # zero-sized work is fine everywhere
    e = np.zeros((0, 8), np.uint16)
    r = B.playouts(e, 16)
    assert len(r["wins_p0"]) == 0
    r = B.playouts(s, 0)
    assert r["wins_p0"].shape == (4,)
    p = B.perft(e, 3)
    assert p["nodes"].tolist() == [0, 0, 0, 0]
    assert B.perft(s, 0)["nodes"].tolist() == [4]
    assert len(B.canon_dedup(e)[0]) == 0 and len(B.bfs_level(e)[0]) == 0
    # full boards and positions without moves
    full = O.make_roots(1000, seed=3, min_plies=12, span=3)
    g, o = B.playouts(full, 32, SEED, per_playout=True), O.playouts(full, 32, SEED, per_playout=True)
    assert np.array_equal(g["outcome"], o["outcome"]) and np.array_equal(g["plies"], o["plies"])
    for k, v in O.perft(full, 3).items():
        assert np.array_equal(B.perft(full, 3)[k], v), k


# ----------------------------------------------------------------------------- eval-guided rollouts
def test_evaluation_and_guided_rollouts():
    g = _npz("guided_rollouts.npz")
    st = g["states"]
    n = len(st)
    f0, f1 = B.eval_features(st, np.zeros(n, np.uint8)), B.eval_features(st, np.ones(n, np.uint8))
    assert np.array_equal(f0, g["features_p0"]) and np.array_equal(f1, g["features_p1"])       # integers: exact
    # evaluate = weights @ features in float64; tolerance 1e-12 absolute (summation order of a 6-term dot product)
    assert np.abs(B.evaluate(st, np.zeros(n, np.uint8), g["fitted_weights"]) - g["evaluate_p0_fitted"]).max() < 1e-12
    big = O.make_roots(20000, seed=123, min_plies=1, span=13)
    pl = (np.arange(len(big)) % 2).astype(np.uint8)
    assert np.array_equal(B.eval_features(big, pl), O.features(big, pl))
    from quantik_b200 import _native as N
    # both scoring modes of the kernel (0 = the children of all lanes dealt out over the warp, 1 = each lane scans
    # its own children) must play the same games
    for scoring in (0, 1):
        N.set_option("guided_scoring", scoring)
        e = np.zeros((1, 8), np.uint16)
        for name, wkey, eps in (("seeded_eps02", "seeded_weights", 0.2), ("fitted_eps02", "fitted_weights", 0.2),
                                ("fitted_eps0", "fitted_weights", 0.0), ("seeded_eps1", "seeded_weights", 1.0)):
            r = B.guided_playouts(e, len(g[name + "_outcome"]), SEED, eps, g[wkey], root_index_base=3, per_playout=True)
            assert np.array_equal(r["outcome"][0], g[name + "_outcome"]) and np.array_equal(r["plies"][0], g[name + "_plies"]), name
        r = B.guided_playouts(g["roots"], 24, SEED, 0.2, g["fitted_weights"], first_rollout=500, root_index_base=40, per_playout=True)
        assert np.array_equal(r["outcome"], g["roots_outcome"]) and np.array_equal(r["plies"], g["roots_plies"])
        r = B.guided_playouts(e, 80, SEED, 0.2, g["fitted_weights"], root_index_base=3, max_plies=9, per_playout=True)
        assert np.array_equal(r["outcome"][0], g["cut9_outcome"]) and np.array_equal(r["plies"][0], g["cut9_plies"])
        # larger sample against the oracle, both weight sets, per playout
        roots = O.make_roots(64, seed=9)
        for w, eps in ((g["fitted_weights"], 0.2), (g["seeded_weights"], 0.05)):
            a, b = B.guided_playouts(roots, 128, SEED, eps, w, per_playout=True), O.guided_playouts(roots, 128, SEED, eps, w, per_playout=True)
            assert np.array_equal(a["outcome"], b["outcome"]) and np.array_equal(a["plies"], b["plies"])
            assert np.array_equal(a["wins_p0"], b["wins_p0"]) and np.array_equal(a["wins_p1"], b["wins_p1"])
    N.set_option("guided_scoring", 0)
    with pytest.raises(ValueError):
        B.guided_playouts(e, 4, epsilon=1.5)                                                    # mcts.py:70-76


def test_exact_endgame_solver():
    g = _npz("solver.npz")
    oc, pl, act = B.solve(g["states"])
    assert np.array_equal(oc.astype(float) * (float(g["win"]) - pl), g["score"])          # MinimaxEngine.solve(...).score
    # MinimaxEngine.solve(...).best_move: one of the optimal moves (which one depends on the reference's iterative-deepening
    # history); the set of ALL optimal moves must contain it, and qk_solve's own move is the lowest member of that set
    score, optimal = B.solve_moves(g["states"])
    ref_a = g["best_action"].astype(np.uint64)
    assert np.all((optimal >> ref_a) & np.uint64(1))
    live = act != 255
    assert np.array_equal(act[live], np.array([(int(m) & -int(m)).bit_length() - 1 for m in optimal[live]], np.uint8))
    assert np.array_equal(score[np.arange(len(act))[live], act[live]], (oc.astype(np.int16) * (64 - pl.astype(np.int16)))[live])
    o = O.solve(g["states"])
    assert np.array_equal(oc, o[0]) and np.array_equal(pl, o[1]) and np.array_equal(act, o[2])
    # 3000 positions with 4..9 empty squares + finished / blocked / invalid ones against the oracle negamax
    roots = np.concatenate([O.make_roots(3000, seed=321, min_plies=7, span=6),
                            np.array([Q.bb_from_qfen("A..C/bbd./CD.A/.adB"), Q.bb_from_qfen("AbCd/..../..../...."),
                                      (1, 1, 0, 0, 0, 0, 0, 0)], np.uint16)])
    a, b = B.solve(roots), O.solve(roots)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert set(np.unique(a[0])) == {-1, 1} and a[1].max() >= 5
    # consistency: playing the reported move flips the outcome and shortens the game by one ply
    live = a[2] != 255
    kids = B.apply_moves(roots[live], a[2][live])
    k = B.solve(kids)
    assert np.array_equal(-k[0], a[0][live]) and np.array_equal(k[1] + 1, a[1][live])


# ----------------------------------------------------------------------------- device buffers, engines
def test_device_tensor_path_matches_host_path():
    import torch
    s = _random_positions(50_000, seed=9)
    t = torch.from_numpy(s.view(np.int16)).cuda()
    mask, tm, st = B.movegen_masks(t)
    hm, htm, hst = B.movegen_masks(s)
    assert np.array_equal(mask.cpu().numpy().view(np.uint64), hm) and np.array_equal(tm.cpu().numpy(), htm)
    assert np.array_equal(B.canonical_payloads(t).cpu().numpy().view(np.uint16), B.canonical_payloads(s))
    roots = torch.from_numpy(O.make_roots(512).view(np.int16)).cuda()
    with torch.cuda.stream(torch.cuda.Stream()):
        r = B.playouts(roots, 512, SEED)
        torch.cuda.current_stream().synchronize()
    h = B.playouts(roots.cpu().numpy().view(np.uint16), 512, SEED)
    assert np.array_equal(r["wins_p0"].cpu().numpy().view(np.uint32), h["wins_p0"])
    assert np.array_equal(r["wins_p1"].cpu().numpy().view(np.uint32), h["wins_p1"])


def test_engine_plugins():
    class FakeState:
        def __init__(self, bb):
            self.bb = tuple(int(x) for x in bb)

    roots = O.make_roots(5, seed=11)
    ev = Q.GpuRolloutEvaluator(rollouts=512, seed=SEED)
    vals = [ev(FakeState(r)) for r in roots]                 # BeamSearchConfig.evaluator protocol
    o = O.playouts(roots, 512, SEED)
    want = (o["wins_p0"].astype(float) - o["wins_p1"]) / 512
    assert np.allclose(vals, want) and all(-1.0 <= v <= 1.0 for v in vals)
    ev2 = Q.GpuRolloutEvaluator(rollouts=512, seed=SEED)
    assert np.allclose(ev2.evaluate_many(roots), want)
    stats = {"rollouts": 0}
    f = Q.gpu_default_evaluate(Q.GpuRolloutEvaluator(rollouts=8, seed=SEED))
    v = f(None, FakeState(roots[0]), 64, stats)
    assert stats["rollouts"] == 64 and -1.0 <= v <= 1.0      # tests/test_beam_search.py:1023-1035

    class Node:
        def __init__(self, flags, depth):
            self.flags, self.depth = flags, depth

    class Tree:
        def __init__(self, node, bb):
            self._n, self._s = node, FakeState(bb)

        def get_node(self, _):
            return self._n

        def get_state(self, _):
            return self._s

    class Cfg:
        max_depth = 16

    class Engine:
        config = Cfg()

    sim = Q.gpu_simulate(rollouts_per_leaf=256, seed=SEED)
    eng = Engine()
    eng.tree = Tree(Node(0x01 | 0x04, 3), roots[0])
    assert sim(eng, 0) == 1.0                                 # terminal flag short-cut (mcts.py:396-403)
    eng.tree = Tree(Node(0, 16), roots[0])
    assert sim(eng, 0) == 0.0                                 # depth >= max_depth -> 0.0 (mcts.py:412,437)
    eng.tree = Tree(Node(0, 0), np.zeros(8, np.uint16))
    o = O.playouts(np.zeros((1, 8), np.uint16), 256, SEED, root_index_base=1, max_plies=16)
    assert abs(sim(eng, 0) - (float(o["wins_p0"][0]) - float(o["wins_p1"][0])) / 256) < 1e-12

    class EvalCfg:                                            # EvalConfig stand-in (evaluation.py:46-62)
        weights = np.array(Q.SEEDED_WEIGHTS)

    Cfg.rollout_eval_config, Cfg.rollout_epsilon = EvalCfg(), 0.1
    o = O.guided_playouts(np.zeros((1, 8), np.uint16), 256, SEED, 0.1, Q.SEEDED_WEIGHTS, root_index_base=2, max_plies=16)
    assert abs(sim(eng, 0) - (float(o["wins_p0"][0]) - float(o["wins_p1"][0])) / 256) < 1e-12
    Cfg.rollout_eval_config = None


def test_exchange_step_functions_at_world_size_one_and_two_gpus():
    """quantik_b200.dist's global dedup and partitioned BFS.  With one process they must equal
    the plain calls; when the box has >= 2 GPUs the real NCCL all-to-all runs under torchrun (scripts/multi_gpu_bfs.py)."""
    import subprocess
    import sys
    import torch
    from quantik_b200 import dist as D
    rows = D.sharded_symmetry_table(7, device=0)
    assert [list(map(int, r)) for r in B.symmetry_table(7)] == rows
    pos = _random_positions(4096)
    pos = pos[O.validate(pos)[1] == 0]
    own, mult, n_global = D.sharded_canon_dedup(pos, None, device=0, sort_output=True)
    u, um = B.canon_dedup(pos)
    assert n_global == len(u) and np.array_equal(np.asarray(own), u) and np.array_equal(np.asarray(mult), um)
    # owners partition the classes evenly enough to be useful
    owners = D.owner_rank(u, 8).numpy()
    assert np.bincount(owners, minlength=8).min() > len(u) // 16
    # the NCCL path's owner function (torch) is the library's (qk_payload_owner): both exchanges partition alike
    for w in (1, 2, 3, 8, 16):
        assert np.array_equal(D.owner_rank(u, w).numpy(), B.payload_owner(u, w).astype(np.int64)), w
    if torch.cuda.device_count() >= 2:
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        import socket
        with socket.socket() as sock:                     # a free rendezvous port
            sock.bind(("127.0.0.1", 0))
            port = sock.getsockname()[1]
        out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                              "--master-addr", "127.0.0.1", "--master-port", str(port),
                              os.path.join(root, "scripts", "multi_gpu_bfs.py"), "9", "18"],
                             capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        rec = json.loads(out.stdout.strip().splitlines()[-1])
        assert rec["bfs_matches_single_gpu"] and rec["dedup_matches_single_gpu"] and rec["n_gpus"] == 2
        assert rec["fused_bfs_matches_single_gpu"] and rec["fused_dedup_matches_single_gpu"]      # peer-memory path


@pytest.mark.parametrize("dedup_mode", [1, 2])
def test_fused_exchange_with_emulated_ranks_on_one_gpu(options, dedup_mode):
    """qk_canon_dedup_scatter / qk_bfs_level_scatter / qk_inbox_merge with W ranks played by ONE GPU (every inbox
    is a local buffer): the owners' shares must be disjoint and add up to the single-table result; a BFS run
    through the inboxes must give the published table.  The real peer-memory run is scripts/multi_gpu_bfs.py.
    dedup_mode 1 / 2: the local merge in one table (delivered from the table) / in two-level form (delivered from
    its dense list of classes)."""
    import torch
    options("dedup_mode", dedup_mode)
    W, sps = 3, 40_000
    nbytes = B.inbox_bytes(W, sps)
    bufs = [torch.empty(nbytes, dtype=torch.uint8, device="cuda:0") for _ in range(W)]
    ptrs = [b.data_ptr() for b in bufs]
    pos = _random_positions(120_000, seed=99)[:60_000]
    pos = np.concatenate([pos, np.full((5, 8), 0xFFFF, np.uint16)])          # the payload kept outside the tables
    mult = np.random.default_rng(5).integers(1, 1000, len(pos)).astype(np.uint64)
    sent = np.zeros((W, W), np.uint64)
    for r, (s_, m_) in enumerate(zip(np.array_split(pos, W), np.array_split(mult, W))):
        sent[r] = B.canon_dedup_scatter(s_, m_, ptrs, r, sps)
    want_u, want_m = B.canon_dedup(pos, mult)
    got_u, got_m = [], []
    for r in range(W):
        received = int(bufs[r][: 8 * W].view(torch.int64).sum().item())
        assert received == int(sent[:, r].sum())
        u, m = B.inbox_merge(ptrs[r], W, sps, max(received, 1), sort_output=True)
        got_u.append(u.cpu().numpy().view(np.uint16)); got_m.append(m.cpu().numpy().view(np.uint64))
    assert min(len(u) for u in got_u) > len(want_u) // (2 * W)                 # owners share the classes evenly
    allu, allm = np.concatenate(got_u), np.concatenate(got_m)
    assert len(allu) == len(want_u)                                            # disjoint shares
    order = np.lexsort(allu.astype("<u2").view(np.uint8).reshape(-1, 16).T[::-1])
    assert np.array_equal(allu[order], want_u) and np.array_equal(allm[order], want_m)
    with pytest.raises(Q.QuantikB200Error):                                     # region too small -> QK_E_NOMEM
        B.canon_dedup_scatter(pos, mult, ptrs, 0, 100)

    # BFS through the inboxes: depth 1..6 from the empty board with W emulated ranks
    sps = 400_000
    bufs = [torch.empty(B.inbox_bytes(W, sps), dtype=torch.uint8, device="cuda:0") for _ in range(W)]
    ptrs = [b.data_ptr() for b in bufs]
    shards = [(np.zeros((1 if r == 0 else 0, 8), np.uint16), np.ones(1 if r == 0 else 0, np.uint64)) for r in range(W)]
    rows = []
    for depth in range(1, 7):
        tot = np.zeros(5, np.int64)
        for r, (f, m) in enumerate(shards):
            st, _ = B.bfs_level_scatter(f, m, ptrs, r, sps, local_cap=max(4096, 40 * len(f)))
            tot += np.array(list(st.values()), np.int64)
        shards = []
        for r in range(W):
            received = int(bufs[r][: 8 * W].view(torch.int64).sum().item())
            u, m = B.inbox_merge(ptrs[r], W, sps, max(received, 1))
            shards.append((u.cpu().numpy().view(np.uint16), m.cpu().numpy().view(np.uint64)))
        rows.append([int(tot[0]), sum(len(u) for u, _ in shards), int(tot[1] + tot[4]), int(tot[2] + tot[3]),
                     int(sum(int(m.sum()) for _, m in shards))])
    assert rows == [list(map(int, r)) for r in B.symmetry_table(6)]


def test_single_root_kernel_matches_general_kernel_and_oracle(options):
    """Config 1 (one root, tallies only) runs a specialised kernel; it must give the general kernel's and the oracle's tallies."""
    roots = [np.zeros((1, 8), np.uint16), O.make_roots(3, seed=5, min_plies=6, span=3)[1:2],
             np.array([[1, 2, 4, 8, 0, 0, 0, 0]], np.uint16)]           # the last one is already won (row of A B C D)
    for root in roots:
        for n, first in ((1, 0), (1000, 0), (300_001, 12345), (1 << 22, 1 << 40)):
            tallies = {}
            for variant in ("auto", "fresh", "hoist", "general", "pred5", "pred6"):
                options("playout_variant", {"auto": 0, "hoist": 1, "general": 2, "pred5": 3, "pred6": 4, "fresh": 6}[variant])
                r = B.playouts(root, n, SEED, first_rollout=first)
                tallies[variant] = (int(r["wins_p0"][0]), int(r["wins_p1"][0]), int(r["draws"][0]))
                assert sum(tallies[variant]) == n
            assert len(set(tallies.values())) == 1, (n, first, tallies)
            if n <= 300_001:
                o = O.playouts(root, n, SEED, first)
                assert tallies["auto"] == (int(o["wins_p0"][0]), int(o["wins_p1"][0]), int(o["draws"][0]))
    options("playout_variant", 0)


def test_many_roots_tally_kernel_matches_general_kernel_and_oracle(options):
    """Tallies-only playouts of many roots run the predicated kernel with (root, slice) work items; every root's tallies
    must equal the general kernel's and the oracle's -- odd slice sizes, decided roots, fewer items than lanes."""
    roots = O.make_roots(3000, seed=11, min_plies=0, span=14)
    roots[5] = np.array([1, 2, 4, 8, 0, 0, 0, 0], np.uint16)                  # already won
    roots[6] = np.array([1, 0, 0, 0, 0, 0, 0, 0], np.uint16)                  # ordinary, P1 to move
    for n_roots, rollouts, first, grab in ((3000, 100, 0, 0), (3000, 100, 7, 7), (37, 1000, 1 << 33, 0), (512, 16, 3, 64), (3000, 8, 0, 3)):
        got = {}
        for variant in ("general", "auto", "pred5", "pred6"):
            options("playout_variant", {"auto": 0, "general": 2, "pred5": 3, "pred6": 4}[variant])
            options("playout_grab", grab)
            r = B.playouts(roots[:n_roots], rollouts, SEED, first_rollout=first, root_index_base=5)
            got[variant] = np.stack([np.asarray(r[k].cpu() if hasattr(r[k], "cpu") else r[k]) for k in ("wins_p0", "wins_p1", "draws")])
            assert (got[variant].sum(axis=0) == rollouts).all()
        for variant in ("auto", "pred5", "pred6"):
            assert np.array_equal(got[variant], got["general"]), (n_roots, rollouts, variant)
        o = O.playouts(roots[:n_roots], rollouts, SEED, first, root_index_base=5)
        assert np.array_equal(got["auto"], np.stack([o["wins_p0"], o["wins_p1"], o["draws"]]))
    # the cut-off and the per-playout outputs run the same kernel with a few more instructions per ply
    for n_roots, rollouts, mp, grab in ((300, 37, -1, 0), (300, 37, 0, 0), (300, 37, 3, 5), (300, 64, 7, 0), (40, 500, 16, 0), (300, 1, 9, 0)):
        got = {}
        for variant in ("general", "auto"):
            options("playout_variant", {"auto": 0, "general": 2}[variant])
            options("playout_grab", grab)
            r = B.playouts(roots[:n_roots], rollouts, SEED, max_plies=mp, first_rollout=3, root_index_base=1, per_playout=True)
            got[variant] = [np.asarray(r[k].cpu() if hasattr(r[k], "cpu") else r[k]) for k in ("wins_p0", "wins_p1", "draws", "outcome", "plies")]
        o = O.playouts(roots[:n_roots], rollouts, SEED, 3, root_index_base=1, max_plies=mp, per_playout=True)
        for i, k in enumerate(("wins_p0", "wins_p1", "draws", "outcome", "plies")):
            assert np.array_equal(got["auto"][i], got["general"][i]), (n_roots, rollouts, mp, k)
            assert np.array_equal(got["auto"][i].reshape(np.asarray(o[k]).shape), o[k]), (n_roots, rollouts, mp, k, "oracle")
    options("playout_variant", 0)
    options("playout_grab", 0)
