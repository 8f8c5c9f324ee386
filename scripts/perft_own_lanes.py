#!/usr/bin/env python3
"""k_perft_tile: from how many giving lanes on should every giver evaluate the common part of its child list itself?"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantik-core-py_b200"))
import quantik_b200 as Q  # noqa: E402

e = np.zeros((1, 8), np.uint16)
Q.perft(e, 4)
for depth in (6, 7, 8):
    for own in (33, 64 + 24, 128 + 24, 128 + 8):
        Q._native.set_option("perft_own_lanes", own)
        Q.perft(e, depth)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            p = Q.perft(e, depth)
            ts.append(time.perf_counter() - t0)
        print(f"depth {depth} own_lanes {own:2d}: {min(ts)*1e3:9.3f} ms  nodes[{depth}]={int(p['nodes'][depth])}", flush=True)
d4 = np.load(os.path.join(ROOT, "tests", "golden", "depth4_canonical.npz"))
for own in (33, 64 + 24, 128 + 24, 128 + 8):
    Q._native.set_option("perft_own_lanes", own)
    Q.perft(d4["states"], 6, d4["mult"])
    t0 = time.perf_counter()
    p = Q.perft(d4["states"], 6, d4["mult"])
    print(f"depth-4 classes + 6, own_lanes {own:3d}: {(time.perf_counter() - t0) * 1e3:9.3f} ms  nodes[4]={int(p['nodes'][4])}", flush=True)
