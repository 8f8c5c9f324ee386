#!/usr/bin/env python3
"""Round-2 A/B timings on one GPU: canonical BFS, canonical dedup, guided playouts, perft frontier thresholds."""
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
from quantik_b200 import _native as N  # noqa: E402

out = {}
what = sys.argv[1:] or ["bfs", "dedup", "guided", "perft"]


def best(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return min(ts)


if "bfs" in what:
    out["bfs16_s"] = best(lambda: Q.symmetry_table(16))
if "dedup" in what:
    pos = Q.random_roots(1 << 24, min_plies=3, span=6, as_torch=True)
    for mode, name in ((1, "one_level"), (2, "two_level")):
        N.set_option("dedup_mode", mode)
        t = best(lambda: Q.canon_dedup(pos, None, sort_output=False), 5)
        out[f"dedup_2p24_{name}_keys_per_s"] = (1 << 24) / t
    N.set_option("dedup_mode", 0)
    del pos
if "guided" in what:
    e = np.zeros((1, 8), np.uint16)
    t = best(lambda: Q.guided_playouts(e, 1 << 24))
    out["guided_playouts_per_s"] = (1 << 24) / t
    roots = Q.random_roots(4096, as_torch=True)
    t = best(lambda: Q.guided_playouts(roots, 1024))
    out["guided_midgame_playouts_per_s"] = (1 << 22) / t
if "perft" in what:
    d4 = np.load(os.path.join(ROOT, "tests", "golden", "depth4_canonical.npz"))
    idx = D.partition_by_cost(Q.perft_probe(d4["states"]), 8)[0]
    s, m = np.ascontiguousarray(d4["states"][idx]), np.ascontiguousarray(d4["mult"][idx])
    e = np.zeros((1, 8), np.uint16)
    res = {}
    for mi in (0, 100_000, 300_000, 600_000, 1_200_000, 5_000_000):
        N.set_option("perft_min_items", mi)
        res[mi] = {"share8_depth6_ms": 1e3 * best(lambda: Q.perft(s, 6, m), 5), "whole_depth6_ms": 1e3 * best(lambda: Q.perft(d4["states"], 6, d4["mult"])),
                   "empty_depth6_ms": 1e3 * best(lambda: Q.perft(e, 6), 5), "empty_depth7_ms": 1e3 * best(lambda: Q.perft(e, 7))}
    N.set_option("perft_min_items", 0)
    out["perft_min_items"] = res
print(json.dumps(out))
