#!/usr/bin/env python3
"""One rank's share of the sharded depth-10 enumeration (world 8, largest-first dealing), run three times: for an ncu
launch list (where does the per-rank fixed cost go?) and a host-side timeline of the call."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantik-core-py_b200"))
import quantik_b200 as Q  # noqa: E402
from quantik_b200 import dist as D  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 6
d4 = np.load(os.path.join(ROOT, "tests", "golden", "depth4_canonical.npz"))
idx = D.partition_by_cost(Q.perft_probe(d4["states"]), world)[0]
s, m = np.ascontiguousarray(d4["states"][idx]), np.ascontiguousarray(d4["mult"][idx])
for i in range(int(sys.argv[3]) if len(sys.argv) > 3 else 3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    Q.perft(s, depth, m)
    print(f"run {i}: {1e3 * (time.perf_counter() - t0):.3f} ms for {len(s)} roots")
