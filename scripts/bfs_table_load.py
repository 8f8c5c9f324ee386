#!/usr/bin/env python3
"""Canonical BFS to depth 12 with a given merge-table size (slots per expected class x 10): ncu launch-list target
for the question "is k_bfs_expand_insert bound by DRAM sectors touched per child?"."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantik-core-py_b200"))
import quantik_b200 as Q  # noqa: E402
from quantik_b200 import _native as N  # noqa: E402

x10 = int(sys.argv[1]) if len(sys.argv) > 1 else 15
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 12
N.init(0)
N.set_option("table_slots_x10", x10)
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rows = Q.symmetry_table(depth)
    torch.cuda.synchronize()
    print(f"slots x10 = {x10}: bfs({depth}) {time.perf_counter() - t0:.4f} s, classes {rows[-1][1]}")
