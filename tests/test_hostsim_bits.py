"""The kernels' __host__ __device__ rule code (csrc/qk_bits.cuh, qk_playout_core.cuh,
qk_perft_core.cuh) compiled for the CPU and checked against the reference-generated golden
vectors and the oracle.  Catches bit-trick errors without a GPU; the GPU tier re-checks the
same functions through the C-ABI."""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
G = os.path.join(HERE, "golden")


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.fixture(scope="module")
def hs():
    subprocess.check_call(["sh", os.path.join(HERE, "hostsim", "build.sh")])
    lib = ctypes.CDLL(os.path.join(HERE, "hostsim", "_build", "libqk_hostsim.so"))
    lib.hs_d4.restype = ctypes.c_uint16
    return lib


def enc(t):
    t = t.astype(np.uint32)
    return np.ascontiguousarray(t[:, 0] | (t[:, 1] << 3) | (t[:, 2] << 4) | (t[:, 3] << 6) | (t[:, 4] << 8) | (t[:, 5] << 10))


def test_rules_primitives(hs):
    d = np.load(os.path.join(G, "states_reference.npz"))
    s = np.ascontiguousarray(d["states"])
    n = len(s)
    mask, tm, st = np.zeros(n, np.uint64), np.zeros(n, np.uint8), np.zeros(n, np.uint8)
    hs.hs_movegen(P(s), ctypes.c_size_t(n), P(mask), P(tm), P(st))
    assert np.array_equal(mask, d["mask"]) and np.array_equal(tm, d["to_move"]) and np.array_equal(st, d["v_code"])
    w = np.zeros(n, np.uint8)
    hs.hs_win(P(s), ctypes.c_size_t(n), P(w))
    assert np.array_equal(w, d["win"])
    for cs, key, trk in ((0, "canon192", "tr192"), (1, "canon384", "tr384")):
        out, tr = np.zeros_like(s), np.zeros(n, np.uint32)
        hs.hs_canonical(P(s), ctypes.c_size_t(n), cs, P(out), P(tr))
        assert np.array_equal(out, d[key])
        assert np.array_equal(tr, enc(d[trk]))
    o = np.zeros(n, np.uint8)
    hs.hs_orbit(P(s), ctypes.c_size_t(n), P(o))
    assert np.array_equal(o, d["orbit"])


def test_d4_maps_all_masks(hs):
    for d4 in range(8):
        for v in range(0, 65536, 3):
            assert hs.hs_d4(v, d4) == O.lib().qo_permute16(v, d4)


def test_apply_symmetry_images(hs):
    g = np.load(os.path.join(G, "symmetry_images.npz"))
    s = np.ascontiguousarray(g["states"])
    out = np.zeros_like(s)
    hs.hs_apply_symmetry(P(s), P(enc(g["transform"])), ctypes.c_size_t(len(s)), P(out))
    assert np.array_equal(out, g["image"])


def test_select_kth_bit(hs):
    rng = random.Random(7)
    for _ in range(5000):
        m = rng.getrandbits(64) & rng.getrandbits(64)
        if not m:
            continue
        k = rng.randrange(bin(m).count("1"))
        assert hs.hs_select(m & 0xFFFFFFFF, m >> 32, k) == [b for b in range(64) if m >> b & 1][k]


def test_playout_step_vs_reference_rules(hs):
    p = np.load(os.path.join(G, "playouts_philox.npz"))
    seed = int(p["seed"])
    e = np.zeros((1, 8), np.uint16)

    def run(roots, R, first, base, mp):
        oc, pl = np.zeros((len(roots), R), np.int8), np.zeros((len(roots), R), np.uint8)
        hs.hs_playouts(P(roots), ctypes.c_size_t(len(roots)), ctypes.c_uint32(R), ctypes.c_uint64(seed),
                       ctypes.c_uint64(first), ctypes.c_uint32(base), ctypes.c_int(mp), P(oc), P(pl))
        return oc, pl

    oc, pl = run(e, len(p["empty_outcome"]), 0, 0, -1)
    assert np.array_equal(oc[0], p["empty_outcome"]) and np.array_equal(pl[0], p["empty_plies"])
    r = np.ascontiguousarray(p["roots"])
    oc, pl = run(r, p["roots_outcome"].shape[1], int(p["roots_first_rollout"]), int(p["roots_index_base"]), -1)
    assert np.array_equal(oc, p["roots_outcome"]) and np.array_equal(pl, p["roots_plies"])
    for mp in (16, 12, 5, 0):
        oc, pl = run(e, len(p[f"cut{mp}_outcome"]), 0, 0, mp)
        assert np.array_equal(oc[0], p[f"cut{mp}_outcome"]) and np.array_equal(pl[0], p[f"cut{mp}_plies"])


def test_predicated_ply_model_vs_reference_rules(hs):
    """The ply k_playout_pred runs in PTX, restated in C++ on the same tables (move-table image, 16-bit byte table, the
    "blocked" 65th entry, the sign test): the reference's outcomes under the shared stream, playout for playout."""
    p = np.load(os.path.join(G, "playouts_philox.npz"))
    seed = int(p["seed"])

    def run(roots, R, first, base):
        oc, sl = np.zeros((len(roots), R), np.int8), np.zeros((len(roots), R), np.uint8)
        hs.hs_playouts_pred(P(roots), ctypes.c_size_t(len(roots)), ctypes.c_uint32(R), ctypes.c_uint64(seed),
                            ctypes.c_uint64(first), ctypes.c_uint32(base), P(oc), P(sl))
        return oc, sl

    oc, sl = run(np.zeros((1, 8), np.uint16), len(p["empty_outcome"]), 0, 0)
    assert np.array_equal(oc[0], p["empty_outcome"])
    extra = sl[0].astype(int) - p["empty_plies"].astype(int)            # a blocked mover takes a slot but makes no move
    assert set(np.unique(extra)) <= {0, 1} and (extra == 1).any()
    r = np.ascontiguousarray(p["roots"])
    oc, sl = run(r, p["roots_outcome"].shape[1], int(p["roots_first_rollout"]), int(p["roots_index_base"]))
    assert np.array_equal(oc, p["roots_outcome"])
    extra = sl.astype(int) - p["roots_plies"].astype(int)
    assert set(np.unique(extra)) <= {0, 1}


def test_leaf_parent_bulk_counts(hs):
    d = np.load(os.path.join(G, "states_reference.npz"))
    nv = int(d["n_valid"])
    s = np.ascontiguousarray(d["states"][:nv][d["win"][:nv] == 0][:600])
    n = len(s)
    mv, mo, wi = np.zeros(n, np.uint8), np.zeros(n, np.uint8), np.zeros(n, np.uint8)
    hs.hs_leaf_parent(P(s), ctypes.c_size_t(n), P(mv), P(mo), P(wi))
    for i in range(n):
        r = O.perft(s[i:i + 1], 1)
        assert int(r["nodes"][1]) == mo[i]
        assert int(r["wins_p0"][1]) + int(r["wins_p1"][1]) == wi[i]
        assert (int(r["wins_p0"][1]) > 0) <= (mv[i] == 0) and (int(r["wins_p1"][1]) > 0) <= (mv[i] == 1)


def test_leaf_parent_from_parent_data(hs):
    """The perft tile kernel never materialises a leaf parent: its legal squares and line presence are derived from
    its parent's -- in the generic form (dealt stage) and in the shape-specialised form with hoisted masks (own lists).
    Same counts as the direct evaluation, for every non-winning child of reachable positions."""
    d = np.load(os.path.join(G, "states_reference.npz"))
    nv = int(d["n_valid"])
    s = np.concatenate([d["states"][:nv][d["win"][:nv] == 0], O.make_roots(4000, seed=99, min_plies=0, span=14)])
    s = np.ascontiguousarray(s[(O.win_status(s) == 0) & (O.validate(s)[1] == 0)])
    checked, bad = ctypes.c_uint64(0), ctypes.c_uint64(0)
    hs.hs_leaf_from_parent(P(s), ctypes.c_size_t(len(s)), ctypes.byref(checked), ctypes.byref(bad))
    assert checked.value > 20_000 and bad.value == 0


def test_eval_features_and_guided_policy(hs):
    g = np.load(os.path.join(G, "guided_rollouts.npz"))
    st = np.ascontiguousarray(g["states"])
    n = len(st)
    for pl, key in ((0, "features_p0"), (1, "features_p1")):
        out = np.zeros((n, 6))
        hs.hs_features(P(st), P(np.full(n, pl, np.uint8)), ctypes.c_size_t(n), P(out))
        assert np.array_equal(out, g[key])
    e = np.zeros((1, 8), np.uint16)
    seed = int(np.load(os.path.join(G, "playouts_philox.npz"))["seed"])
    for name, wkey, eps in (("seeded_eps02", "seeded_weights", 0.2), ("fitted_eps02", "fitted_weights", 0.2),
                            ("fitted_eps0", "fitted_weights", 0.0), ("seeded_eps1", "seeded_weights", 1.0)):
        m = len(g[name + "_outcome"])
        oc, pl, w = np.zeros(m, np.int8), np.zeros(m, np.uint8), np.ascontiguousarray(g[wkey])
        hs.hs_guided(P(e), ctypes.c_size_t(1), ctypes.c_uint32(m), ctypes.c_uint64(seed), ctypes.c_uint64(0), ctypes.c_uint32(3),
                     ctypes.c_int(-1), ctypes.c_double(eps), P(w), P(oc), P(pl))
        assert np.array_equal(oc, g[name + "_outcome"]) and np.array_equal(pl, g[name + "_plies"]), name
    # the packed-field evaluation (12 lines at once) against the oracle's line-by-line loop on 20 000 reachable positions,
    # and guided playouts from mid-game roots against the oracle's
    big = np.ascontiguousarray(O.make_roots(20000, seed=77, min_plies=0, span=15))
    for pl in (0, 1):
        out = np.zeros((len(big), 6))
        hs.hs_features(P(big), P(np.full(len(big), pl, np.uint8)), ctypes.c_size_t(len(big)), P(out))
        assert np.array_equal(out, O.features(big, np.full(len(big), pl, np.uint8)))
    roots = np.ascontiguousarray(O.make_roots(24, seed=5, min_plies=2, span=9))
    w = np.ascontiguousarray(g["fitted_weights"])
    for eps, mp, nr in ((0.2, -1, 40), (0.0, -1, 40), (0.5, 4, 40), (0.3, -1, 3), (1.0, -1, 40)):      # 3 rollouts: no greedy line
        oc, pl = np.zeros((24, nr), np.int8), np.zeros((24, nr), np.uint8)
        hs.hs_guided(P(roots), ctypes.c_size_t(24), ctypes.c_uint32(nr), ctypes.c_uint64(seed), ctypes.c_uint64(7), ctypes.c_uint32(0),
                     ctypes.c_int(mp), ctypes.c_double(eps), P(w), P(oc), P(pl))
        o = O.guided_playouts(roots, nr, seed, eps, w, 7, 0, mp, True)
        assert np.array_equal(oc, o["outcome"]) and np.array_equal(pl, o["plies"]), (eps, mp, nr)


def test_child_features_from_parent_data(hs):
    """k_playout_guided scores a child without building it: its line fields come from the parent's plus one table word,
    and a threat line counts for a player exactly when the completing placement is among his legal moves, which follow
    from the parent's (qk_eval_core.cuh: eval_features_child).  Same six features as evaluating the child itself
    (evaluation.py:130-194), for every child of 24 000 reachable positions."""
    s = np.ascontiguousarray(O.make_roots(24000, seed=321, min_plies=0, span=15))
    checked, bad = ctypes.c_uint64(0), ctypes.c_uint64(0)
    hs.hs_child_features(P(s), ctypes.c_size_t(len(s)), ctypes.byref(checked), ctypes.byref(bad))
    assert checked.value > 200_000 and bad.value == 0, (checked.value, bad.value)


def test_fuzz_random_boards_vs_oracle(hs):
    """Hypothesis-style sweep (tests/test_core.py:126-181 of the reference): arbitrary 8 x 16-bit boards -- reachable,
    sparse garbage and dense garbage -- through the kernels' rule code vs the oracle's literal loops."""
    rng = np.random.default_rng(2026)
    reach = O.make_roots(6000, seed=99, min_plies=0, span=14)
    sparse = (rng.integers(0, 65536, (6000, 8)) & rng.integers(0, 65536, (6000, 8)) & rng.integers(0, 65536, (6000, 8))
              & rng.integers(0, 65536, (6000, 8))).astype(np.uint16)
    dense = rng.integers(0, 65536, (3000, 8)).astype(np.uint16)
    mutated = reach[:3000].copy()
    mutated[np.arange(3000), rng.integers(0, 8, 3000)] |= (1 << rng.integers(0, 16, 3000)).astype(np.uint16)
    s = np.ascontiguousarray(np.concatenate([reach, sparse, dense, mutated]))
    n = len(s)
    mask, tm, st = np.zeros(n, np.uint64), np.zeros(n, np.uint8), np.zeros(n, np.uint8)
    hs.hs_movegen(P(s), ctypes.c_size_t(n), P(mask), P(tm), P(st))
    om, ot, os_ = O.movegen_masks(s)
    assert np.array_equal(mask, om) and np.array_equal(tm, ot) and np.array_equal(st, os_)
    assert set(np.unique(st)) >= {0, 1, 2, 4, 5}
    w = np.zeros(n, np.uint8)
    hs.hs_win(P(s), ctypes.c_size_t(n), P(w))
    assert np.array_equal(w, O.win_status(s))
    for cs in (0, 1):
        out, tr = np.zeros_like(s), np.zeros(n, np.uint32)
        hs.hs_canonical(P(s), ctypes.c_size_t(n), cs, P(out), P(tr))
        oc, otr = O.canonical(s, bool(cs), True)
        assert np.array_equal(out, oc)
        code = (otr[:, 0] | (otr[:, 1] << 3) | (otr[:, 3] << 4) | (otr[:, 4] << 6) | (otr[:, 5] << 8) | (otr[:, 6] << 10)).astype(np.uint32)
        assert np.array_equal(tr, code)
    o = np.zeros(n, np.uint8)
    hs.hs_orbit(P(s), ctypes.c_size_t(n), P(o))
    assert np.array_equal(o, O.orbit_size(s))
