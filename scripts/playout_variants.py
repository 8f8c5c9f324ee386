#!/usr/bin/env python3
"""Config-1 throughput of the playout kernel variants (qk_set_option "playout_variant") and a tally cross-check."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantik-core-py_b200"))
sys.path.insert(0, ROOT)
import quantik_b200 as Q  # noqa: E402
import quantik_b200.batch as B  # noqa: E402
from oracle import oracle as O  # noqa: E402

VARIANTS = {"auto": 0, "hoist": 1, "general": 2, "pred5": 3, "pred6": 4, "pred4": 5, "fresh": 6}
names = [a for a in sys.argv[1:] if not a.startswith("grab=")] or ["fresh", "pred5", "pred6", "auto"]
grabs = [int(a[5:]) for a in sys.argv[1:] if a.startswith("grab=")] or [0]

roots = [np.zeros((1, 8), np.uint16), O.make_roots(3, seed=5, min_plies=6, span=3)[1:2], O.make_roots(9, seed=7, min_plies=9, span=3)[4:5],
         np.array([[1, 2, 4, 8, 0, 0, 0, 0]], np.uint16)]
bad = 0
for ri, root in enumerate(roots):
    for n, first in ((1, 0), (1000, 0), (300_001, 12345), (1 << 22, 1 << 40)):
        want = None
        for v in ["general"] + names:
            Q._native.set_option("playout_variant", VARIANTS[v])
            r = B.playouts(root, n, 20260101, first_rollout=first)
            t = (int(r["wins_p0"][0]), int(r["wins_p1"][0]), int(r["draws"][0]))
            if want is None:
                want = t
            elif t != want:
                bad += 1
                print("MISMATCH root", ri, "n", n, "first", first, v, t, "general", want, flush=True)
print("tally cross-check:", "OK" if not bad else f"{bad} mismatches", flush=True)

# config 4: 65 536 mid-game roots x 1 024 rollouts
roots4 = B.random_roots(65536, seed=4, min_plies=4, span=8, as_torch=True)
if roots4 is not None:
    for variant in ["general"] + [v for v in names if v != "fresh"]:
        Q._native.set_option("playout_variant", VARIANTS[variant])
        Q._native.set_option("playout_grab", 0)
        B.playouts(roots4, 64)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(3):
            r = B.playouts(roots4, 1024, first_rollout=i * 1024)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        print(f"config 4 {variant:8s} {65536 * 1024 / dt:.4e} playouts/s  p0 {float(r['wins_p0'].sum()) / (65536 * 1024):.5f}", flush=True)

root = torch.zeros((1, 8), dtype=torch.int16, device="cuda:0")
n = 1 << 30
for variant, grab in [(v, g) for g in grabs for v in names] * 2:
    Q._native.set_option("playout_variant", VARIANTS[variant])
    Q._native.set_option("playout_grab", grab)
    Q.playouts(root, 1 << 24)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(3):
        r = Q.playouts(root, n, first_rollout=i * n)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f"{variant:8s} grab {grab:3d} {n / dt:.4e} playouts/s  p0 {int(r['wins_p0'][0]) / n:.5f}", flush=True)
