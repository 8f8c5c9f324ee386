#!/usr/bin/env python3
"""BASELINE config 5 on one GPU: raw depth-first enumeration below the 10 946 canonical depth-4
roots (weighted by multiplicity) to the end of the game, cross-checked against the canonical BFS
(two independent kernels: k_perft_dfs vs k_bfs_expand_insert)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantik-core-py_b200"))
import quantik_b200 as Q  # noqa: E402

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 12
d4 = np.load(os.path.join(ROOT, "tests", "golden", "depth4_canonical.npz"))
t0 = time.perf_counter()
p = Q.perft(d4["states"], depth, d4["mult"])
dt = time.perf_counter() - t0
nodes = int(p["nodes"][1:].sum())
print(f"perft depth-4 roots +{depth}: {dt:.2f} s, {nodes:.3e} nodes, {nodes / dt:.3e} nodes/s")
print("checking against the canonical BFS ...", flush=True)
rows = Q.symmetry_table(4 + depth)
ok = True
for r in range(1, depth + 1):
    d = 4 + r
    want = rows[d - 1]
    got = [int(p["nodes"][r]), int(p["wins_p0"][r]) + int(p["blocked_p1_loses"][r - 1]),
           int(p["wins_p1"][r]) + int(p["blocked_p0_loses"][r - 1]),
           int(p["nodes"][r]) - int(p["wins_p0"][r]) - int(p["wins_p1"][r])]
    match = got == [want[0], want[2], want[3], want[4]]
    ok &= match
    print(d, got, "OK" if match else f"MISMATCH vs {want}")
print("ALL MATCH" if ok else "MISMATCH")
