#!/usr/bin/env python3
"""Config-1 throughput of the playout kernel variants (qk_set_option "playout_variant": fresh | hoist | general)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantik-core-py_b200"))
import quantik_b200 as Q  # noqa: E402

root = torch.zeros((1, 8), dtype=torch.int16, device="cuda:0")
n = 1 << 30
for variant in ("general", "hoist", "fresh", "general", "hoist", "fresh"):
    Q._native.set_option("playout_variant", {"fresh": 0, "hoist": 1, "general": 2}[variant])
    Q.playouts(root, 1 << 24)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(3):
        r = Q.playouts(root, n, first_rollout=i * n)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f"{variant:8s} {n / dt:.4e} playouts/s  p0 {int(r['wins_p0'][0]) / n:.5f}", flush=True)
