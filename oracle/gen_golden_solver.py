#!/usr/bin/env python3
"""Golden vectors for the exact endgame solver: MinimaxEngine.solve (minimax.py:182-204) of the reference on
late-game positions.  Run in the build container: python oracle/gen_golden_solver.py   (~2 min)"""
import os
import sys
import time

import numpy as np

REF = os.environ.get("QUANTIK_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(REF, "src"))
sys.path.insert(0, os.path.join(HERE, ".."))
from quantik_core import State  # noqa: E402
from quantik_core.minimax import MinimaxConfig, MinimaxEngine  # noqa: E402
import oracle as O  # noqa: E402

roots = np.concatenate([O.make_roots(150, seed=2024, min_plies=9, span=4), O.make_roots(30, seed=2025, min_plies=8, span=1)])
eng = MinimaxEngine(MinimaxConfig())
scores, moves, t0 = [], [], time.time()
for i, bb in enumerate(roots):
    r = eng.solve(State(tuple(int(x) for x in bb)))
    scores.append(r.score)
    moves.append(r.best_move.shape * 16 + r.best_move.position)
    if i % 30 == 0:
        print(i, r.score, time.time() - t0, flush=True)
np.savez_compressed(os.path.join(HERE, "..", "tests", "golden", "solver.npz"), states=roots,
                    score=np.array(scores, np.float64), best_action=np.array(moves, np.uint8), win=np.array(10000.0))
print("done", time.time() - t0)
