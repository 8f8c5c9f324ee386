#!/usr/bin/env python3
"""Strong scaling of the sharded enumeration (BASELINE config 5 shortened: the 10 946 weighted depth-4 classes, 6 plies
deeper) predicted on ONE GPU: every rank's share of quantik_b200.dist.partition_by_cost is timed alone; the job time at
world N is the slowest share (+ one 680-byte all-reduce).  Prints shares for round-robin and largest-first dealing."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantik-core-py_b200"))
import quantik_b200 as Q  # noqa: E402
from quantik_b200 import dist as D  # noqa: E402

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 6
d4 = np.load(os.path.join(ROOT, "tests", "golden", "depth4_canonical.npz"))
states, mult = d4["states"], d4["mult"]


def timed(s, m, reps=5):
    Q.perft(s, depth, m)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        Q.perft(s, depth, m)
        best = min(best, time.perf_counter() - t0)
    return best


t0 = time.perf_counter()
cost = Q.perft_probe(states)
t_probe = time.perf_counter() - t0
t0 = time.perf_counter()
cost = Q.perft_probe(states)
t_probe = time.perf_counter() - t0
whole = timed(states, mult)
out = {"depth": depth, "whole_s": whole, "probe_s": t_probe, "worlds": {}}
for world in (2, 4, 8):
    lpt = D.partition_by_cost(cost, world)
    rr = [np.arange(r, len(states), world) for r in range(world)]
    res = {}
    for name, parts in (("largest_first", lpt), ("round_robin", rr)):
        ts = [timed(np.ascontiguousarray(states[i]), np.ascontiguousarray(mult[i]), 3) for i in parts]
        res[name] = {"max_s": max(ts), "min_s": min(ts), "efficiency": whole / world / max(ts),
                     "cost_imbalance": float(max(cost[i].sum() for i in parts) / (cost.sum() / world))}
    out["worlds"][world] = res
print(json.dumps(out))
