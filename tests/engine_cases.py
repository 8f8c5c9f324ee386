"""Checks of the level-batched beam engine and the batched-leaf MCTS engine (quantik_b200.engines) against the
reference's OWN engines.  Shared by the CPU tier (tests/test_adapters_with_reference.py: the oracle stands in for the
kernels, reference from /root/reference) and the GPU tier (tests/test_gpu_engines.py: the real kernels, reference from
the recipe-installed oracle/_ref)."""
import numpy as np


def _leaf_tuple(leaf):
    return (tuple((m.player, m.shape, m.position) for m in leaf.moves), float(leaf.value), leaf.depth, leaf.is_terminal,
            leaf.multiplicity)


def _node_rows(tree):
    rows = []
    for nid in range(tree.storage.node_count):
        n = tree.get_node(nid)
        rows.append((int(n.parent_id), int(n.flags), int(n.visit_count), float(n.best_value), int(n.depth), int(n.player_turn)))
    return rows


def beam_identical_with_deterministic_evaluator(Q, qc, qfens=("..../..../..../....", "A.bC/..../d..B/...a", "Ab../..C./.d../...B")):
    """With the same deterministic evaluator the batched engine must reproduce the stock search EXACTLY: same best
    line, same leaves in the same order, same statistics, same tree.  (_expand_frontier, beam_search.py:411-546, is the
    part that moved to the GPU here; a custom evaluator without evaluate_many keeps the per-candidate scoring.)"""
    from quantik_core import beam_search as bs
    from quantik_core.core import State

    def evaluator(state):                        # deterministic, cheap, spreads over [-1, 1], not symmetric in the players
        bb = state.bb
        h = 0
        for i, w in enumerate(bb):
            h = (h * 1000003 + (int(w) + 1) * (i + 7)) & 0xFFFFFFFF
        return (h % 2001) / 1000.0 - 1.0

    Batched = Q.batched_beam_engine(bs)
    for qfen in qfens:
        for cfg_kw in (dict(beam_width=6, max_depth=4), dict(beam_width=40, max_depth=3, beam_schedule=[40, 10, 5]),
                       dict(beam_width=3, max_depth=16)):
            res = []
            for cls in (bs.BeamSearchEngine, Batched):
                eng = cls(bs.BeamSearchConfig(evaluator=evaluator, random_seed=1, **cfg_kw))
                r = eng.search(State.from_qfen(qfen))
                stats = {k: v for k, v in r.stats.items() if k != "memory_usage"}
                c = eng._counters
                res.append((_leaf_tuple(r.best_leaf), [_leaf_tuple(x) for x in r.terminal_leaves],
                            [_leaf_tuple(x) for x in r.frontier_leaves], r.reached_terminal, r.max_depth_reached, stats, r.root_player,
                            _node_rows(eng.tree), (c.expanded_nodes, c.generated_nodes, c.terminal_hits, c.canonical_dedup_hits),
                            eng._root_dedup_hits))
            assert res[0] == res[1], (qfen, cfg_kw)
    return True


def beam_default_evaluator_bookkeeping(Q, qc, playouts_check=None):
    """The built-in rollout evaluator, one launch per level: the counters _default_evaluate keeps (beam_search.py:614-622)
    equal the stock engine's, every value is a mean of `rollouts` +-1 results, and the values are the Philox tallies of
    the level's candidates (playouts_check(roots, rollouts, seed, first_rollout) -> wins_p0, wins_p1 recomputes them)."""
    from quantik_core import beam_search as bs
    from quantik_core.core import State
    Batched = Q.batched_beam_engine(bs, seed=77)
    cfg = dict(beam_width=5, max_depth=3, rollouts_per_candidate=16, random_seed=3)
    stock = bs.BeamSearchEngine(bs.BeamSearchConfig(**cfg))
    rs = stock.search(State.empty())
    eng = Batched(bs.BeamSearchConfig(**cfg))
    rb = eng.search(State.empty())
    # depth 1: 64 moves -> 3 classes, all evaluated in both engines; deeper levels depend on which survive
    assert rb.stats["evaluations"] * 16 == rb.stats["rollouts"] and rs.stats["evaluations"] * 16 == rs.stats["rollouts"]
    assert rb.stats["candidates_generated"] >= 64 and rb.max_depth_reached == 3 and len(rb.frontier_leaves) == 5
    for leaf in rb.frontier_leaves:
        assert abs(leaf.value * 16 - round(leaf.value * 16)) < 1e-9 and -1.0 <= leaf.value <= 1.0
    assert eng.gpu_launches <= 6 * 3                      # a fixed number of launches per level, not one per candidate
    if playouts_check is not None:
        # level 1 from the empty board: the candidates are the first move of each of the 3 classes, in move order, and
        # candidate i is root i of the level's launch (seed, first_rollout 0)
        eng2 = Batched(bs.BeamSearchConfig(beam_width=64, max_depth=1, rollouts_per_candidate=64, random_seed=3))
        leaves = sorted(eng2.search(State.empty()).frontier_leaves, key=lambda leaf: (leaf.moves[0].shape, leaf.moves[0].position))
        assert len(leaves) == 3
        roots = np.array([qc.move.apply_move((0,) * 8, leaf.moves[0]) for leaf in leaves], np.uint16)
        w0, w1 = playouts_check(roots, 64, 77, 0)
        assert [round(leaf.value * 64) for leaf in leaves] == [int(a) - int(b) for a, b in zip(w0, w1)]
    return True


def mcts_batched_invariants(Q, qc, guided=False):
    """Batched-leaf MCTS: the playout budget is spent exactly (in units of rollouts_per_leaf), every playout is booked on
    the whole path (root visits = playouts, P0 + P1 wins + cut-off draws = visits), the returned move is legal, and the
    statistics of the root's children add up to the root's."""
    from quantik_core import mcts as mc
    from quantik_core.core import State
    from quantik_core.move import generate_legal_moves_list
    kw = {}
    if guided:
        from quantik_core.evaluation import EvalConfig
        kw = dict(rollout_eval_config=EvalConfig(), rollout_epsilon=0.2)
    Batched = Q.batched_mcts_engine(mc, leaves_per_wave=16, rollouts_per_leaf=32, seed=5)
    for qfen, budget in (("..../..../..../....", 2048), ("A.bC/..../d..B/...a", 1024), ("C.A./...c/D.B./b.dc", 512)):
        eng = Batched(mc.MCTSConfig(max_iterations=budget, random_seed=42, **kw))
        state = State.from_qfen(qfen)
        move, value = eng.search(state)
        assert move in generate_legal_moves_list(state.bb) and 0.0 <= value <= 1.0
        assert eng.iterations_performed == budget and eng.waves == budget // (16 * 32)
        assert eng.gpu_launches <= 2 * 17 * eng.waves      # one launch per distinct remaining depth per wave, not per leaf
        root = eng.tree.get_node(eng.root_id)
        assert int(root.visit_count) == budget
        # (a game decided by the 16th piece is a cut-off draw in the reference: its loop stops at depth == max_depth
        # before looking at the board again, mcts.py:412-437)
        assert 0.8 * budget < int(root.win_count_p0) + int(root.win_count_p1) <= budget
        kids = [eng.tree.get_node(c) for c in eng.tree.get_children(eng.root_id)]
        assert sum(int(k.visit_count) for k in kids) <= budget
        assert abs(float(root.best_value) - (int(root.win_count_p0) - int(root.win_count_p1)) / budget) < 1e-6
    # a cut-off (max_depth below the game length) books draws as the reference does: visits without a win
    eng = Batched(mc.MCTSConfig(max_iterations=512, max_depth=6, random_seed=1, **kw))
    eng.search(State.empty())
    root = eng.tree.get_node(eng.root_id)
    assert int(root.visit_count) == 512 and int(root.win_count_p0) + int(root.win_count_p1) < 512
    return True
