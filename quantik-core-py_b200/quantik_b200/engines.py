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
        if self._next_root + len(roots) > 1 << 32:      # root_index_base is a u32 Philox counter word
            raise OverflowError("GpuRolloutEvaluator: 2^32 roots evaluated with one seed; start a new evaluator (new seed)")
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
        # root_index_base is a u32 Philox counter word: the call number's low half goes there, its high half moves
        # the 64-bit rollout ids on, so the streams never repeat
        call = next(counter)
        base, first = call & 0xFFFFFFFF, (call >> 32) * rollouts_per_leaf
        if cfg is not None:
            r = B.guided_playouts(np.array([bb], np.uint16), rollouts_per_leaf, seed=seed, max_plies=left,
                                  epsilon=float(getattr(self.config, "rollout_epsilon", 0.2)),
                                  weights=np.asarray(cfg.weights, np.float64), first_rollout=first, root_index_base=base,
                                  device=device)
        else:
            r = B.playouts(np.array([bb], np.uint16), rollouts_per_leaf, seed=seed, max_plies=left,
                           first_rollout=first, root_index_base=base, device=device)
        return (float(r["wins_p0"][0]) - float(r["wins_p1"][0])) / rollouts_per_leaf

    return _simulate


# ----------------------------------------------------------------------------- level-batched beam search
_KEY_HEADER = b"\x01\x02"       # State.canonical_key = VERSION, FLAG_CANON + payload (core.py:171-180)


def batched_beam_engine(beam_module, device: int = 0, seed: int = B.SEED_DEFAULT):
    """-> a subclass of the reference's ``BeamSearchEngine`` (pass ``quantik_core.beam_search``) in which every depth
    level costs a fixed number of kernel launches instead of one Python rule call per move and per rollout:

    * ``_expand_frontier`` (beam_search.py:411-461, ``_expand_moves`` :463-546): ONE ``qk_movegen_masks`` over the
      whole frontier, ONE ``qk_apply_moves`` for all its children, ONE ``qk_win_status``, ONE ``qk_canonical`` for the
      children that did not finish the game; the dictionary / tree bookkeeping that follows is the reference's, in the
      reference's order (first-encountered move kept, multiplicities summed, same counters).
    * ``_score_and_prune`` (:549-590; the candidate loop :565-569): ONE ``qk_playouts`` launch evaluates every
      candidate of the level with the built-in random-rollout evaluator (``rollouts`` per candidate,
      ``stats["rollouts"]`` kept as ``_default_evaluate`` does, :614-622).  A custom ``config.evaluator`` keeps the
      reference's one-call-per-candidate path (its cost model is its own, :601-612) unless it offers
      ``evaluate_many`` (``GpuRolloutEvaluator`` does), which is then called once per level.

    The search logic (frontier, pruning order, tree, telemetry) stays the reference's; only the data-parallel work
    moved.  Rollout values come from the Philox stream (seed, level candidate index, rollout), not from MT19937."""
    Base = beam_module.BeamSearchEngine
    BeamLeaf, Move, State = beam_module.BeamLeaf, beam_module.Move, beam_module.State
    F_P0, F_P1 = beam_module.NODE_FLAG_WINNING_P0, beam_module.NODE_FLAG_WINNING_P1

    class BatchedBeamSearchEngine(Base):
        gpu_device, gpu_seed = int(device), int(seed)

        def __init__(self, config):
            super().__init__(config)
            self._stream_pos = 0            # rollouts consumed so far: every level draws fresh Philox rollout ids
            self.gpu_launches = 0

        def _expand_frontier(self, frontier, depth, stats, terminal_leaves):
            candidates = {}
            if not frontier:
                return candidates
            dev = self.gpu_device
            parents = np.array([tuple(e[1]) for e in frontier], dtype=np.uint16).reshape(-1, 8)
            masks, movers, _status = B.movegen_masks(parents, device=dev)
            bits = np.unpackbits(np.ascontiguousarray(masks).view(np.uint8).reshape(-1, 8), axis=1, bitorder="little")
            pi, act = np.nonzero(bits)                          # parent-major, ascending action = the reference's move order
            kids = wins = keys = None
            if len(pi):
                kids = B.apply_moves(parents[pi], act.astype(np.uint8), device=dev)
                wins = np.asarray(B.win_status(kids, device=dev))
                open_idx = np.flatnonzero(wins == 0)
                keys = {}
                if len(open_idx):
                    raw = np.ascontiguousarray(B.canonical_payloads(kids[open_idx], device=dev)).astype("<u2").tobytes()
                    keys = {int(c): _KEY_HEADER + raw[16 * i:16 * i + 16] for i, c in enumerate(open_idx)}
                self.gpu_launches += 4
            else:
                self.gpu_launches += 1
            first = np.searchsorted(pi, np.arange(len(frontier) + 1))
            for f, (node_id, _bb, moves, _value, multiplicity) in enumerate(frontier):
                self._counters.expanded_nodes += 1
                lo, hi = int(first[f]), int(first[f + 1])
                mover = int(movers[f])
                if lo == hi:                                     # the mover cannot move: the other player wins (:424-447)
                    value = 1.0 if mover == 1 else -1.0
                    self._mark_terminal(node_id, F_P0 if mover == 1 else F_P1, value)
                    terminal_leaves.append(BeamLeaf(moves=moves, value=value, depth=depth - 1, is_terminal=True,
                                                    multiplicity=multiplicity))
                    self._counters.terminal_hits += 1
                    continue
                stats["candidates_generated"] += hi - lo
                for c in range(lo, hi):
                    self._counters.generated_nodes += 1
                    a = int(act[c])
                    move = Move(mover, a >> 4, a & 15)
                    child_bb = tuple(int(x) for x in kids[c])
                    line = moves + (move,)
                    w = int(wins[c])
                    if w:                                        # a finished child is a leaf of its own (:491-526)
                        value = 1.0 if w == 1 else -1.0
                        before = self.tree.storage.node_count
                        child_id = self.tree.add_child_node(node_id, State(child_bb), multiplicity=multiplicity)
                        self._mark_terminal(child_id, F_P0 if w == 1 else F_P1, value)
                        if self.tree.storage.node_count > before:
                            stats["nodes_inserted"] += 1
                        terminal_leaves.append(BeamLeaf(moves=line, value=value, depth=depth, is_terminal=True,
                                                        multiplicity=multiplicity))
                        self._counters.terminal_hits += 1
                        continue
                    key = keys[c]
                    seen = candidates.get(key)
                    if seen is not None:                         # symmetric sibling: only the weight accumulates (:528-545)
                        stats["candidates_deduped"] += 1
                        self._counters.canonical_dedup_hits += 1
                        if depth == 1:
                            self._root_dedup_hits += 1
                        candidates[key] = seen[:4] + (seen[4] + multiplicity,)
                    else:
                        candidates[key] = (node_id, child_bb, line, mover, multiplicity)
            return candidates

        def _level_values(self, boards, rollouts, stats):
            """raw evaluator values (player-0 view) of every candidate of a level"""
            ev = self.config.evaluator
            if ev is not None and not hasattr(ev, "evaluate_many"):
                return [self._evaluate(State(bb), rollouts, stats) for bb in boards]
            if ev is not None:
                vals = ev.evaluate_many(boards)
            else:
                roots = np.array(boards, dtype=np.uint16).reshape(-1, 8)
                r = B.playouts(roots, int(rollouts), seed=self.gpu_seed, first_rollout=self._stream_pos, device=self.gpu_device)
                self._stream_pos += int(rollouts)
                stats["rollouts"] += int(rollouts) * len(boards)
                vals = (np.asarray(r["wins_p0"], np.float64) - np.asarray(r["wins_p1"], np.float64)) / float(rollouts)
            self.gpu_launches += 2
            return [max(-1.0, min(1.0, float(v))) for v in vals]

        def _score_and_prune(self, candidates, stats, beam_width, rollouts):
            if not candidates:
                return []
            entries = list(candidates.items())
            raw = self._level_values([e[1][1] for e in entries], rollouts, stats)
            stats["evaluations"] += len(entries)
            order = sorted(range(len(entries)), key=lambda i: (-(raw[i] if entries[i][1][3] == 0 else -raw[i]), i))
            keep = order[:beam_width]
            stats["nodes_pruned"] += max(0, len(order) - len(keep))
            frontier = []
            for i in keep:
                parent_id, bb, moves, _mover, multiplicity = entries[i][1]
                before = self.tree.storage.node_count
                child_id = self.tree.add_child_node(parent_id, State(bb), multiplicity=multiplicity)
                node = self.tree.get_node(child_id)
                node.best_value = np.float32(raw[i])
                node.visit_count = np.uint32(node.visit_count + 1)
                self.tree.storage.store_node(child_id, node)
                if self.tree.storage.node_count > before:
                    stats["nodes_inserted"] += 1
                frontier.append((child_id, bb, moves, raw[i], multiplicity))
            return frontier

    BatchedBeamSearchEngine.__name__ = "BatchedBeamSearchEngine"
    return BatchedBeamSearchEngine


# ----------------------------------------------------------------------------- batched-leaf MCTS
def batched_mcts_engine(mcts_module, leaves_per_wave: int = 64, rollouts_per_leaf: int = 256, device: int = 0,
                        seed: int = B.SEED_DEFAULT):
    """-> a subclass of the reference's ``MCTSEngine`` (pass ``quantik_core.mcts``) whose ``search`` (mcts.py:95-148)
    runs in WAVES: ``leaves_per_wave`` descents (the reference's own ``_select`` / ``_expand``, :150-325) are made
    with their visits booked at once -- a virtual loss, so consecutive descents spread over the tree -- then ONE
    ``qk_playouts`` launch plays ``rollouts_per_leaf`` playouts from every collected leaf (one launch per distinct
    remaining depth ``config.max_depth - node.depth``, the cut-off of ``_simulate``, :412,436-437), and the tallies
    are propagated to the root as ``_backpropagate`` (:439-476) does, ``rollouts_per_leaf`` results at a time.
    ``config.max_iterations`` stays the PLAYOUT budget: ``iterations_performed`` counts simulated games, so a stock
    and a batched engine with the same config spend the same number of playouts.  With ``rollout_eval_config`` the
    playouts use the eval-guided policy (``qk_playouts_guided``)."""
    Base = mcts_module.MCTSEngine
    F_TERM, F_P0, F_P1 = mcts_module.NODE_FLAG_TERMINAL, mcts_module.NODE_FLAG_WINNING_P0, mcts_module.NODE_FLAG_WINNING_P1
    K, R = int(leaves_per_wave), int(rollouts_per_leaf)
    if K <= 0 or R <= 0:
        raise ValueError("leaves_per_wave and rollouts_per_leaf must be positive")

    class BatchedMCTSEngine(Base):
        gpu_device, gpu_seed = int(device), int(seed)

        def __init__(self, config):
            super().__init__(config)
            self._stream_pos = 0
            self.gpu_launches = 0
            self.waves = 0

        # -- the two halves of _backpropagate (:439-476): visits when the leaf is chosen, results when they arrive
        def _path(self, node_id):
            path = []
            cur = node_id
            while cur is not None:
                path.append(cur)
                if cur == self.root_id:
                    break
                cur = int(self.tree.get_node(cur).parent_id)
            return path

        def _book_visits(self, path, n):
            for nid in path:
                node = self.tree.get_node(nid)
                node.visit_count = np.uint32(int(node.visit_count) + n)
                self.tree.storage.store_node(nid, node)

        def _book_results(self, path, w0, w1):
            for nid in path:
                node = self.tree.get_node(nid)
                node.win_count_p0 = np.uint32(int(node.win_count_p0) + w0)
                node.win_count_p1 = np.uint32(int(node.win_count_p1) + w1)
                visits = int(node.visit_count)
                # the running average of +1 / -1 / 0 results (:463-467) is (P0 wins - P1 wins) / visits
                node.best_value = np.float32((int(node.win_count_p0) - int(node.win_count_p1)) / visits if visits else 0.0)
                self.tree.storage.store_node(nid, node)

        def _simulate_wave(self, leaves):
            """(wins_p0, wins_p1) out of R playouts for every leaf; terminal leaves return their known value R times
            (:396-403)."""
            out = [None] * len(leaves)
            groups = {}
            for i, nid in enumerate(leaves):
                node = self.tree.get_node(nid)
                flags = int(node.flags)
                if flags & F_TERM:
                    out[i] = (R, 0) if flags & F_P0 else ((0, R) if flags & F_P1 else (0, 0))
                    continue
                left = max(0, int(self.config.max_depth) - int(node.depth))
                groups.setdefault(left, []).append(i)
            cfg = getattr(self.config, "rollout_eval_config", None)
            for left, members in sorted(groups.items()):
                roots = np.array([tuple(self.tree.get_state(leaves[i]).bb) for i in members], dtype=np.uint16).reshape(-1, 8)
                if cfg is not None:
                    r = B.guided_playouts(roots, R, seed=self.gpu_seed, max_plies=left,
                                          epsilon=float(getattr(self.config, "rollout_epsilon", 0.2)),
                                          weights=np.asarray(cfg.weights, np.float64), first_rollout=self._stream_pos,
                                          device=self.gpu_device)
                else:
                    r = B.playouts(roots, R, seed=self.gpu_seed, max_plies=left, first_rollout=self._stream_pos,
                                   device=self.gpu_device)
                self._stream_pos += R
                self.gpu_launches += 2
                for j, i in enumerate(members):
                    out[i] = (int(r["wins_p0"][j]), int(r["wins_p1"][j]))
            return out

        def search(self, initial_state):
            import time as _time
            self.root_id = self.tree.create_root_node(initial_state)
            self.iterations_performed = 0
            self._counters = mcts_module.SearchEventCounters()
            self._max_depth_reached = 0
            self._searched = True
            start = _time.monotonic()
            deadline = start + self.config.time_limit_s if self.config.time_limit_s is not None else None
            budget = int(self.config.max_iterations)
            while self.iterations_performed < budget:
                want = min(K, -(-(budget - self.iterations_performed) // R))
                leaves, paths = [], []
                for _ in range(want):
                    node_id = self._select(self.root_id)
                    if not (int(self.tree.get_node(node_id).flags) & F_TERM):
                        child_id = self._expand(node_id)
                        if child_id is not None:
                            node_id = child_id
                    path = self._path(node_id)
                    self._book_visits(path, R)
                    leaves.append(node_id)
                    paths.append(path)
                for path, (w0, w1) in zip(paths, self._simulate_wave(leaves)):
                    self._book_results(path, w0, w1)
                self.iterations_performed += want * R
                self.waves += 1
                if deadline is not None and _time.monotonic() >= deadline:
                    break
            self._elapsed_ms = int(round((_time.monotonic() - start) * 1000))
            return self._get_best_move()

    BatchedMCTSEngine.__name__ = "BatchedMCTSEngine"
    return BatchedMCTSEngine
