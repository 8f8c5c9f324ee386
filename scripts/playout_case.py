#!/usr/bin/env python3
"""One config-1 launch of a playout kernel variant (for ncu): playout_case.py <variant-number> [log2 playouts]."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantik-core-py_b200"))
import quantik_b200 as Q  # noqa: E402

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 28)
Q._native.set_option("playout_variant", variant)
root = torch.zeros((1, 8), dtype=torch.int16, device="cuda:0")
for i in range(3):
    r = Q.playouts(root, n, first_rollout=i * n)
torch.cuda.synchronize()
print(variant, n, int(r["wins_p0"][0]) / n)
