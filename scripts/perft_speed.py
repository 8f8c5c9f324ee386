#!/usr/bin/env python3
"""perft timing from the empty board for both depth-first kernels (qk_set_option "perft_kernel")."""
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
    for mode in ("tile", "warp", "lane", "auto"):
        Q._native.set_option("perft_kernel", {"auto": 0, "tile": 1, "lane": 2, "warp": 3}[mode])
        if depth == 8 and mode == "warp":
            pass
        Q.perft(e, depth)
        t0 = time.perf_counter()
        p = Q.perft(e, depth)
        dt = time.perf_counter() - t0
        nodes = int(p["nodes"][1:].sum())
        print(f"depth {depth} {mode:5s}: {dt*1e3:9.2f} ms  {nodes/dt:.3e} nodes/s  nodes[{depth}]={int(p['nodes'][depth])}", flush=True)
