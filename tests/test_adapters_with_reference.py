"""CPU tier, build container only: the engine plug-ins (quantik_b200.engines) driven by the REAL reference
engines.  There is no GPU here, so the CUDA batch calls are replaced by the oracle for this test -- what is
being checked is the plug-in protocol (attribute names, flag constants, stats bookkeeping, value ranges),
not the kernels.  Skipped where /root/reference does not exist (the GPU box)."""
import os
import sys

import numpy as np
import pytest

REF = os.environ.get("QUANTIK_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "quantik_core")),
                                reason="reference checkout not available")


@pytest.fixture()
def ref(monkeypatch):
    monkeypatch.syspath_prepend(os.path.join(REF, "src"))
    import oracle as O
    from quantik_b200 import batch as B

    def playouts(roots, rollouts, seed=B.SEED_DEFAULT, max_plies=-1, first_rollout=0, root_index_base=0,
                 per_playout=False, device=None):
        return O.playouts(roots, rollouts, seed, first_rollout, root_index_base, max_plies, per_playout)

    def guided(roots, rollouts, seed=B.SEED_DEFAULT, epsilon=0.2, weights=B.SEEDED_WEIGHTS, max_plies=-1,
               first_rollout=0, root_index_base=0, per_playout=False, device=None):
        return O.guided_playouts(roots, rollouts, seed, epsilon, weights, first_rollout, root_index_base, max_plies, per_playout)

    monkeypatch.setattr(B, "playouts", playouts)
    monkeypatch.setattr(B, "guided_playouts", guided)
    monkeypatch.setattr(B, "movegen_masks", lambda states, device=None: O.movegen_masks(states))
    monkeypatch.setattr(B, "apply_moves", lambda states, actions, device=None: O.apply_moves(states, actions))
    monkeypatch.setattr(B, "win_status", lambda states, device=None: O.win_status(states))
    monkeypatch.setattr(B, "canonical_payloads", lambda states, device=None: O.canonical(states))
    import quantik_core
    return quantik_core


def test_flag_constants_match_reference(ref):
    from quantik_core.memory import compact_tree as ct
    assert (ct.NODE_FLAG_TERMINAL, ct.NODE_FLAG_WINNING_P0, ct.NODE_FLAG_WINNING_P1) == (0x01, 0x04, 0x08)


def test_mcts_engine_with_gpu_simulate(ref, monkeypatch):
    import quantik_b200 as Q
    from quantik_core.core import State
    from quantik_core.evaluation import EvalConfig
    from quantik_core.mcts import MCTSConfig, MCTSEngine
    from quantik_core.move import generate_legal_moves_list
    monkeypatch.setattr(MCTSEngine, "_simulate", Q.gpu_simulate(rollouts_per_leaf=32, seed=7))
    for cfg in (MCTSConfig(max_iterations=40, random_seed=42),
                MCTSConfig(max_iterations=20, random_seed=42, rollout_eval_config=EvalConfig(), rollout_epsilon=0.2)):
        eng = MCTSEngine(cfg)
        state = State.from_qfen("A.bC/..../d..B/...a")
        move, value = eng.search(state)
        assert move in generate_legal_moves_list(state.bb) and -1.0 <= value <= 1.0
        assert eng.iterations_performed == cfg.max_iterations


def test_beam_engine_with_gpu_evaluator(ref, monkeypatch):
    import quantik_b200 as Q
    from quantik_core.beam_search import BeamSearchConfig, BeamSearchEngine
    from quantik_core.core import State
    ev = Q.GpuRolloutEvaluator(rollouts=16, seed=3)
    eng = BeamSearchEngine(BeamSearchConfig(beam_width=4, max_depth=3, evaluator=ev, random_seed=1))
    res = eng.search(State.empty())
    assert res.best_leaf is not None and -1.0 <= res.best_leaf.value <= 1.0 and ev.playouts_run > 0
    assert res.stats["rollouts"] == 0           # a custom evaluator's cost is its own (beam_search.py:601-612)
    # the overridable default evaluator keeps the rollouts counter (tests/test_beam_search.py:1023-1035)
    monkeypatch.setattr(BeamSearchEngine, "_default_evaluate", Q.gpu_default_evaluate(Q.GpuRolloutEvaluator(rollouts=8, seed=3)))
    eng = BeamSearchEngine(BeamSearchConfig(beam_width=4, max_depth=2, rollouts_per_candidate=8, random_seed=1))
    res = eng.search(State.empty())
    assert res.stats["rollouts"] > 0 and res.stats["rollouts"] % 8 == 0


def test_single_state_api_matches_reference_signatures(ref):
    import inspect
    import quantik_b200 as Q
    import quantik_core as R
    from quantik_core import game_utils, move
    for ours, theirs in ((Q.generate_legal_moves, move.generate_legal_moves), (Q.generate_legal_moves_list, move.generate_legal_moves_list),
                         (Q.apply_move, move.apply_move), (Q.validate_move, move.validate_move),
                         (Q.check_game_winner, game_utils.check_game_winner), (Q.has_winning_line, game_utils.has_winning_line)):
        assert list(inspect.signature(ours).parameters) == list(inspect.signature(theirs).parameters), ours.__name__
    for name in ("canonical_key", "canonical_payload", "symmetry_count", "pack", "unpack", "to_qfen", "from_qfen", "empty", "bb",
                 "get_occupied_bb"):
        assert hasattr(Q.State, name) and hasattr(R.State, name)
    assert [int(x) for x in Q.WinStatus] == [int(x) for x in game_utils.WinStatus]
    assert Q.Move(0, 1, 2) == Q.Move(0, 1, 2) and (Q.Move(0, 1, 2).player, Q.Move(0, 1, 2).shape) == (0, 1)


def test_batched_beam_engine_reproduces_stock_search(ref):
    import quantik_b200 as Q
    import engine_cases
    assert engine_cases.beam_identical_with_deterministic_evaluator(Q, ref)


def test_batched_beam_engine_default_evaluator(ref):
    import oracle as O
    import quantik_b200 as Q
    import engine_cases

    def check(roots, rollouts, seed, first):
        o = O.playouts(roots, rollouts, seed, first)
        return o["wins_p0"], o["wins_p1"]

    assert engine_cases.beam_default_evaluator_bookkeeping(Q, ref, check)


def test_batched_mcts_engine(ref):
    import quantik_b200 as Q
    import engine_cases
    assert engine_cases.mcts_batched_invariants(Q, ref)
    assert engine_cases.mcts_batched_invariants(Q, ref, guided=True)
