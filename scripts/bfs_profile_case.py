#!/usr/bin/env python3
"""Profiling case for the merge kernels: the canonical BFS to depth 12 (the largest frontiers) and one
2^24-state canonical dedup.  Used under ncu by profiles/ captures."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantik-core-py_b200"))
import quantik_b200 as Q  # noqa: E402

t0 = time.perf_counter()
rows = Q.symmetry_table(12)
torch.cuda.synchronize()
print("bfs(12)", time.perf_counter() - t0, rows[-1])
roots = Q.random_roots(1 << 24, min_plies=3, span=6)
dev = torch.from_numpy(np.ascontiguousarray(roots).view(np.int16)).cuda()
for _ in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    u, m = Q.canon_dedup(dev, None, sort_output=False)
    torch.cuda.synchronize()
    print("dedup 2^24", time.perf_counter() - t0, int(u.shape[0]))
