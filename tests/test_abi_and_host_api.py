"""CPU tier: the C-ABI library loads and exports every symbol include/quantik_b200.h declares,
fails loudly without a GPU, and the host-side mirror of the reference interface behaves like
the reference where no kernel is needed (argument errors, codecs, transforms)."""
import ctypes
import os
import re

import numpy as np
import pytest

import quantik_b200 as Q
from quantik_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


def header_symbols():
    text = open(os.path.join(ROOT, "include", "quantik_b200.h")).read()
    return re.findall(r"^QK_API\s+[\w \*]+?\b(qk_\w+)\s*\(", text, flags=re.M)


def test_library_exports_every_declared_symbol():
    names = header_symbols()
    assert len(names) >= 30 and len(set(names)) == len(names)
    lib = ctypes.CDLL(N.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(N.PROTOTYPES), "ctypes prototype table out of sync with the header"
    assert N.load().qk_abi_version() == 1


def test_inbox_layout_is_host_arithmetic():
    """qk_inbox_bytes needs no device: header of per-source counts + one (16 B payload + 8 B multiplicity) region per source."""
    from quantik_b200 import batch as B
    assert B.inbox_bytes(1, 1) == 128 + 24
    assert B.inbox_bytes(8, 1000) == 128 + 8 * 1000 * 24
    lib = N.load()
    nb = ctypes.c_size_t(0)
    assert lib.qk_inbox_bytes(17, 10, ctypes.byref(nb)) == -1 and b"world" in lib.qk_last_error()     # QK_MAX_WORLD = 16
    # without an initialised device the scatter entry points refuse like every other compute call
    s = np.zeros((1, 8), np.uint16)
    ptrs = (ctypes.c_void_p * 1)(ctypes.c_void_p(s.ctypes.data))
    if _no_gpu():
        assert lib.qk_canon_dedup_scatter(s.ctypes.data, None, 1, 0, ptrs, 1, 0, 16, None, 0, None) == -3


def test_header_cites_reference_for_every_entry_point():
    text = open(os.path.join(ROOT, "include", "quantik_b200.h")).read()
    for ref in ("move.py:199", "move.py:135", "game_utils.py:123", "symmetry.py:289", "symmetry.py:356",
                "symmetry.py:252", "beam_search.py:614", "mcts.py:385", "game_stats.py:384", "benchmarks/dataset.py:38"):
        assert ref in text, ref


@pytest.mark.skipif(not _no_gpu(), reason="CPU-box behaviour")
def test_no_cpu_fallback_without_gpu():
    lib = N.load()
    assert lib.qk_init(0) == -3    # QK_E_NO_DEVICE
    assert b"no CPU fallback" in lib.qk_last_error()
    with pytest.raises(Q.QuantikB200Error):
        Q.check_game_winner((0,) * 8)
    with pytest.raises(Q.QuantikB200Error):
        Q.playouts(np.zeros((1, 8), np.uint16), 4)
    # uninitialised device: compute entry points refuse instead of computing anything
    out = np.zeros(1, np.uint8)
    s = np.zeros((1, 8), np.uint16)
    assert lib.qk_win_status(s.ctypes.data, 1, out.ctypes.data, 0, None) == -3


def test_argument_errors_match_reference():
    # game_utils.py:289-341 via Move.__post_init__ (move.py:37-39)
    for bad in ((2, 0, 0), (0, 4, 0), (0, 0, 16), (-1, 0, 0)):
        with pytest.raises(ValueError):
            Q.Move(*bad)
    with pytest.raises(ValueError, match="Invalid bitboard data"):     # core.py:30-31
        Q.State((0,) * 7)
    with pytest.raises(ValueError):                                     # bitboard_compact.py:56-58
        Q.State((70000, 0, 0, 0, 0, 0, 0, 0))
    with pytest.raises(ValueError, match="Buffer too small"):          # core.py:60-61
        Q.State.unpack(b"\x01\x00" + b"\x00" * 10)
    with pytest.raises(ValueError, match="Unsupported version"):       # core.py:63-64
        Q.State.unpack(b"\x02\x00" + b"\x00" * 16)
    with pytest.raises(ValueError):                                     # qfen.py:130-131
        Q.bb_from_qfen("..../..../....")
    with pytest.raises(ValueError):                                     # qfen.py:142-145
        Q.bb_from_qfen("X.../..../..../....")
    with pytest.raises(ValueError):                                     # symmetry.py:59-64
        Q.SymmetryTransform(8, False, (0, 1, 2, 3))
    with pytest.raises(ValueError):
        Q.SymmetryTransform(0, False, (0, 1, 2, 2))


def test_pack_qfen_roundtrip_and_layout():
    s = Q.State.from_qfen("A.bC/..../d..B/...a")
    assert s.bb == (1, 2048, 8, 0, 32768, 4, 0, 256)                   # SURVEY 8c, measured with the reference
    assert s.pack() == b"\x01\x00" + np.array(s.bb, "<u2").tobytes() and len(s.pack()) == 18
    assert Q.State.unpack(s.pack(flags=2)) == s
    assert s.to_qfen() == "A.bC/..../d..B/...a"
    assert Q.State.empty().bb == (0,) * 8 and Q.State.empty().get_occupied_bb() == 0
    assert s.get_occupied_bb() == 1 | 2048 | 8 | 32768 | 4 | 256


def test_cbor_roundtrip():
    pytest.importorskip("cbor2")
    s = Q.State.from_qfen("A.bC/..../d..B/...a")
    assert Q.State.from_cbor(s.to_cbor(canon=True, mc=3, meta={"k": 1})) == s
    with pytest.raises(ValueError):
        import cbor2
        Q.State.from_cbor(cbor2.dumps({"v": 2, "bb": b"\x00" * 16}))


def test_transform_codec_and_move_mapping():
    import oracle as O
    g = np.load(os.path.join(ROOT, "tests", "golden", "symmetry_images.npz"))
    for i in range(0, len(g["transform"]), 7):
        d4, cs, *perm = [int(x) for x in g["transform"][i]]
        t = Q.SymmetryTransform(d4, bool(cs), tuple(perm))
        assert Q.decode_transform(t.code()) == (d4, bool(cs), tuple(perm))
        m = Q.Move(*[int(x) for x in g["move_in"][i]])
        m2 = Q.SymmetryHandler.apply_symmetry_to_move(m, t)
        assert [m2.player, m2.shape, m2.position] == g["move_out"][i].tolist()
        inv = t.inverse()
        back = Q.SymmetryHandler.apply_symmetry_to_move(m2, inv)
        assert back == m


def test_symmetry_table_rows_convention():
    p = {"nodes": np.array([1, 10, 50], np.uint64), "wins_p0": np.array([0, 2, 0], np.uint64),
         "wins_p1": np.array([0, 0, 5], np.uint64), "blocked_p0_loses": np.array([0, 0, 0], np.uint64),
         "blocked_p1_loses": np.array([0, 3, 0], np.uint64)}
    assert Q.symmetry_table_rows(p) == [[10, 2, 0, 8], [50, 3, 5, 45]]
