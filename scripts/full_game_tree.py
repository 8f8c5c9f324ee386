#!/usr/bin/env python3
"""Complete canonical enumeration of Quantik (depth 1..16) on one B200: the reference's
GAME_TREE_ANALYSIS.md table (published to depth 8 only) continued to the end of the game with
the fused GPU BFS (qk_bfs_level).  Usage: python scripts/full_game_tree.py [max_depth] [out.json]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantik-core-py_b200"))
import quantik_b200 as Q  # noqa: E402

max_depth = int(sys.argv[1]) if len(sys.argv) > 1 else 16
t0 = time.perf_counter()
rows = Q.symmetry_table(max_depth, verbose=True)
dt = time.perf_counter() - t0
print(f"total {dt:.2f} s")
print("| Depth | Total Legal Moves | Unique Canonical | P0 Wins | P1 Wins | Ongoing |")
print("|---|---|---|---|---|---|")
for d, r in enumerate(rows, 1):
    print(f"| {d} | " + " | ".join(f"{x:,}" for x in r) + " |")
if len(sys.argv) > 2:
    json.dump({"seconds": dt, "rows": rows, "columns": ["total_legal_moves", "unique_canonical", "p0_wins", "p1_wins", "ongoing"]},
              open(sys.argv[2], "w"), indent=1)
