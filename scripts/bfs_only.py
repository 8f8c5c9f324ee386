#!/usr/bin/env python3
"""Complete canonical game tree on one GPU, timed after a warm-up pass (pool sized); ncu launch-list target."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantik-core-py_b200"))
import quantik_b200 as Q  # noqa: E402

Q.symmetry_table(16)
torch.cuda.synchronize()
t0 = time.perf_counter()
rows = Q.symmetry_table(16)
torch.cuda.synchronize()
print("bfs(16) seconds", time.perf_counter() - t0, "states", sum(r[1] for r in rows))
