#!/usr/bin/env python3
"""perft from the empty board, depth 6-8, default kernel choice: median of 5 runs each."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantik-core-py_b200"))
import quantik_b200 as Q  # noqa: E402

e = np.zeros((1, 8), np.uint16)
Q.perft(e, 4)
want = {6: 6241600512, 7: 144052696064, 8: None}
for depth in (6, 7, 8):
    Q.perft(e, depth)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        p = Q.perft(e, depth)
        ts.append(time.perf_counter() - t0)
    dt = sorted(ts)[2]
    nodes = int(p["nodes"][1:].sum())
    ok = want[depth] is None or int(p["nodes"][depth]) == want[depth]
    print(f"depth {depth}: {dt*1e3:9.3f} ms (min {min(ts)*1e3:.3f})  {nodes/dt:.3e} nodes/s  nodes[{depth}]={int(p['nodes'][depth])} {'ok' if ok else 'WRONG'}", flush=True)
