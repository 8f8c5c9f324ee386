#!/usr/bin/env python3
"""BASELINE config 4 (65 536 mid-game roots x 1 024 rollouts, tallies only): ncu target for k_playout_pred_roots."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantik-core-py_b200"))
from quantik_b200 import batch as B  # noqa: E402

roots = B.random_roots(65536, seed=4, min_plies=4, span=8, as_torch=True)
for i in range(3):
    r = B.playouts(roots, 1024, first_rollout=i * 1024)
torch.cuda.synchronize()
print(float(r["wins_p0"].sum()) / (65536 * 1024))
