"""GPU tier: the REAL reference engines (quantik-core itself, from the recipe-installed oracle/_ref -- see
oracle/ref_recipe.py) driven by the REAL kernels through quantik_b200.engines: per-call plug-ins, the level-batched beam
engine and the batched-leaf MCTS engine.  No stand-ins on either side."""
import numpy as np
import pytest

import oracle as O
from oracle import ref_recipe

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_recipe.available(), reason="oracle/_ref missing: run __graft_entry__.build() where /root/reference exists")]


@pytest.fixture(scope="module")
def qc():
    assert ref_recipe.verify(), "oracle/_ref differs from the recorded reference files"
    return ref_recipe.import_reference()


def test_reference_rules_agree_with_kernels_on_this_box(qc):
    """the reference itself, run here, against the kernels: legal moves, winner and canonical key of 300 positions"""
    import quantik_b200 as Q
    from quantik_b200 import batch as B
    from quantik_core.core import State
    from quantik_core.game_utils import check_game_winner
    from quantik_core.move import generate_legal_moves
    roots = O.make_roots(300, seed=31, min_plies=0, span=13)
    masks, movers, _ = B.movegen_masks(roots)
    wins = B.win_status(roots)
    keys = B.canonical_keys(roots)
    for i, r in enumerate(roots):
        bb = tuple(int(x) for x in r)
        player, by_shape = generate_legal_moves(bb)
        want = 0
        for moves in by_shape.values():
            for m in moves:
                want |= 1 << (m.shape * 16 + m.position)
        assert int(masks[i]) == want and int(movers[i]) == int(player)
        assert int(wins[i]) == int(check_game_winner(bb))
        assert keys[i] == State(bb).canonical_key()


def test_stock_engines_with_gpu_plugins(qc):
    import quantik_b200 as Q
    from quantik_core.beam_search import BeamSearchConfig, BeamSearchEngine
    from quantik_core.core import State
    from quantik_core.mcts import MCTSConfig, MCTSEngine
    from quantik_core.move import generate_legal_moves_list
    ev = Q.GpuRolloutEvaluator(rollouts=64, seed=3)
    res = BeamSearchEngine(BeamSearchConfig(beam_width=4, max_depth=3, evaluator=ev, random_seed=1)).search(State.empty())
    assert res.best_leaf is not None and -1.0 <= res.best_leaf.value <= 1.0 and ev.playouts_run > 0

    class Engine(MCTSEngine):
        _simulate = Q.gpu_simulate(rollouts_per_leaf=64, seed=7)

    eng = Engine(MCTSConfig(max_iterations=60, random_seed=42))
    state = State.from_qfen("A.bC/..../d..B/...a")
    move, value = eng.search(state)
    assert move in generate_legal_moves_list(state.bb) and eng.iterations_performed == 60


def test_batched_beam_engine_reproduces_stock_search(qc):
    import quantik_b200 as Q
    import engine_cases
    assert engine_cases.beam_identical_with_deterministic_evaluator(Q, qc)


def test_batched_beam_engine_default_evaluator(qc):
    import quantik_b200 as Q
    import engine_cases

    def check(roots, rollouts, seed, first):
        o = O.playouts(roots, rollouts, seed, first)
        return o["wins_p0"], o["wins_p1"]

    assert engine_cases.beam_default_evaluator_bookkeeping(Q, qc, check)


def test_batched_mcts_engine(qc):
    import quantik_b200 as Q
    import engine_cases
    assert engine_cases.mcts_batched_invariants(Q, qc)
    assert engine_cases.mcts_batched_invariants(Q, qc, guided=True)


def test_batched_engines_search_faster_than_stock_at_equal_budget(qc):
    """benchmarks/adapters.py:160-259 protocol in small: same position, same playout budget, wall time of search()."""
    import time
    import quantik_b200 as Q
    from quantik_core import beam_search as bs, mcts as mc
    from quantik_core.core import State
    state = State.empty()
    budget = 1024
    t0 = time.perf_counter()
    mc.MCTSEngine(mc.MCTSConfig(max_iterations=budget, random_seed=1)).search(state)
    t_stock = time.perf_counter() - t0
    Batched = Q.batched_mcts_engine(mc, leaves_per_wave=16, rollouts_per_leaf=64)
    Batched(mc.MCTSConfig(max_iterations=budget, random_seed=1)).search(state)          # warm-up (context, pool)
    t0 = time.perf_counter()
    eng = Batched(mc.MCTSConfig(max_iterations=budget, random_seed=1))
    eng.search(state)
    t_batched = time.perf_counter() - t0
    assert eng.iterations_performed == budget and t_batched < t_stock / 3, (t_stock, t_batched)
    cfg = dict(beam_width=8, max_depth=3, rollouts_per_candidate=16, random_seed=1)
    t0 = time.perf_counter()
    r_stock = bs.BeamSearchEngine(bs.BeamSearchConfig(**cfg)).search(state)
    t_stock = time.perf_counter() - t0
    BB = Q.batched_beam_engine(bs)
    t0 = time.perf_counter()
    r_b = BB(bs.BeamSearchConfig(**cfg)).search(state)
    t_batched = time.perf_counter() - t0
    assert r_b.stats["rollouts"] > 0 and r_stock.stats["rollouts"] > 0 and t_batched < t_stock / 3, (t_stock, t_batched)
