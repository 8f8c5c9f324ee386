"""Plug-in points the reference's engines expose, backed by batched CUDA playouts.

* ``GpuRolloutEvaluator``  -> ``BeamSearchConfig.evaluator: Callable[[State], float]``
  (beam_search.py:47,601-612) -- mean playout value in [-1, 1] from player 0's view, the
  quantity ``BeamSearchEngine._default_evaluate`` (:614-622) computes on the CPU.
* ``gpu_default_evaluate`` -> replacement for ``BeamSearchEngine._default_evaluate`` that
  also bumps ``stats["rollouts"]`` (pinned by tests/test_beam_search.py:1023-1035).
* ``gpu_simulate``         -> replacement for ``MCTSEngine._simulate`` (mcts.py:385-437):
  K leaf-parallel playouts per call with the reference's depth cut-off semantics.
Nothing here imports the reference; the objects only need ``.bb`` (State) and the
handful of attributes cited.
"""
from __future__ import annotations

import itertools
from typing import Iterable, List, Optional, Sequence

import numpy as np

from . import batch as B


class GpuRolloutEvaluator:
    """``evaluator(state) -> float``: mean of ``rollouts`` uniform-random playouts to a true
    terminal (beam_search.py:624-645), player-0 perspective.  Successive calls consume
    successive root indices of the Philox stream, so a search is reproducible from ``seed``."""

    def __init__(self, rollouts: int = 1024, seed: int = B.SEED_DEFAULT, device: int = 0):
        if rollouts <= 0:
            raise ValueError("rollouts must be positive")     # beam_search.py:63-64
        self.rollouts = int(rollouts)
        self.seed = int(seed)
        self.device = int(device)
        self._next_root = 0
        self.playouts_run = 0

    def evaluate_many(self, bitboards: Sequence[Sequence[int]]) -> np.ndarray:
        """One launch for a whole beam level (the batched form of _score_and_prune's loop,
        beam_search.py:565-569)."""
        roots = np.asarray(bitboards, dtype=np.uint16).reshape(-1, 8)
        r = B.playouts(roots, self.rollouts, seed=self.seed, root_index_base=self._next_root, device=self.device)
        self._next_root += len(roots)
        self.playouts_run += len(roots) * self.rollouts
        w0 = np.asarray(r["wins_p0"], dtype=np.float64)
        w1 = np.asarray(r["wins_p1"], dtype=np.float64)
        return (w0 - w1) / self.rollouts

    def __call__(self, state) -> float:
        bb = state.bb if hasattr(state, "bb") else state
        return float(self.evaluate_many([tuple(bb)])[0])


def gpu_default_evaluate(evaluator: GpuRolloutEvaluator):
    """-> function(self, state, rollouts, stats) usable as BeamSearchEngine._default_evaluate."""

    def _default_evaluate(self, state, rollouts: int, stats) -> float:
        saved = evaluator.rollouts
        evaluator.rollouts = int(rollouts)
        try:
            value = evaluator(state)
        finally:
            evaluator.rollouts = saved
        stats["rollouts"] += rollouts
        return value

    return _default_evaluate


def gpu_simulate(rollouts_per_leaf: int = 256, seed: int = B.SEED_DEFAULT, device: int = 0):
    """-> function(self, node_id) usable as MCTSEngine._simulate (mcts.py:385-437).

    Terminal flags are honoured exactly as the reference does (:396-403); otherwise the leaf is
    evaluated by ``rollouts_per_leaf`` playouts cut off after ``config.max_depth - node.depth``
    moves (:412,436-437) and the mean outcome is returned.  When the engine's config carries
    ``rollout_eval_config`` (mcts.py:47-52) the playouts use the eval-guided epsilon-greedy policy of
    ``_select_rollout_move`` (:345-383) with its weights and ``rollout_epsilon``."""
    counter = itertools.count()
    NODE_FLAG_TERMINAL, NODE_FLAG_WINNING_P0, NODE_FLAG_WINNING_P1 = 0x01, 0x04, 0x08   # memory/compact_tree.py:54-57

    def _simulate(self, node_id: int) -> float:
        node = self.tree.get_node(node_id)
        flags = int(node.flags)
        if flags & NODE_FLAG_TERMINAL:
            if flags & NODE_FLAG_WINNING_P0:
                return 1.0
            if flags & NODE_FLAG_WINNING_P1:
                return -1.0
            return 0.0
        bb = self.tree.get_state(node_id).bb
        left = max(0, int(self.config.max_depth) - int(node.depth))
        cfg = getattr(self.config, "rollout_eval_config", None)
        if cfg is not None:
            r = B.guided_playouts(np.array([bb], np.uint16), rollouts_per_leaf, seed=seed, max_plies=left,
                                  epsilon=float(getattr(self.config, "rollout_epsilon", 0.2)),
                                  weights=np.asarray(cfg.weights, np.float64), root_index_base=next(counter), device=device)
        else:
            r = B.playouts(np.array([bb], np.uint16), rollouts_per_leaf, seed=seed, max_plies=left,
                           root_index_base=next(counter), device=device)
        return (float(r["wins_p0"][0]) - float(r["wins_p1"][0])) / rollouts_per_leaf

    return _simulate
