#!/usr/bin/env python3
"""Canonical BFS to depth 12 with the frontier in table order (the default) or sorted by payload: does the ORDER of the
parents change the cost of k_bfs_expand_insert (duplicates meeting in L2 instead of DRAM)?  ncu launch-list target."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantik-core-py_b200"))
import quantik_b200 as Q  # noqa: E402

sort = bool(int(sys.argv[1])) if len(sys.argv) > 1 else False
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 12
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rows = Q.symmetry_table(depth, sort_output=sort)
    torch.cuda.synchronize()
    print(f"sorted frontier = {sort}: bfs({depth}) {time.perf_counter() - t0:.4f} s, classes {rows[-1][1]}")
