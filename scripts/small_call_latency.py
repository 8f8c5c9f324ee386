#!/usr/bin/env python3
"""Latency of small calls through the public API (BASELINE config 1: 100k playouts from the empty board)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantik-core-py_b200"))
import quantik_b200 as Q  # noqa: E402

e = np.zeros((1, 8), np.uint16)
ed = torch.zeros((1, 8), dtype=torch.int16, device="cuda")
if len(sys.argv) > 1:
    Q._native.init(0)
    Q._native.set_option("call_stage", int(sys.argv[1]))
    print("call_stage option =", sys.argv[1])
for n in (1, 1000, 100_000, 1_000_000, 10_000_000):
    Q.playouts(e, n)
    t0 = time.perf_counter()
    reps = 200 if n <= 1_000_000 else 20
    for i in range(reps):
        r = Q.playouts(e, n, first_rollout=i * n)
    dt = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for i in range(reps):
        r = Q.playouts(ed, n, first_rollout=i * n)
    torch.cuda.synchronize()
    dtd = (time.perf_counter() - t0) / reps
    print(f"n={n:9d}: host-buffer call {dt*1e6:9.1f} us ({n/dt:.3e}/s)   device-buffer call {dtd*1e6:9.1f} us ({n/dtd:.3e}/s)")
s = np.zeros((1, 8), np.uint16)
for name, fn in (("check_game_winner", lambda: Q.check_game_winner((0,) * 8)), ("generate_legal_moves", lambda: Q.generate_legal_moves((0,) * 8)),
                 ("canonical_key", lambda: Q.State.empty().canonical_key())):
    fn()
    t0 = time.perf_counter()
    for _ in range(500):
        fn()
    print(f"{name}: {(time.perf_counter()-t0)/500*1e6:.1f} us per single-state call")
