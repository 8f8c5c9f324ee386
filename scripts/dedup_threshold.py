#!/usr/bin/env python3
"""qk_canon_dedup, one table against the two-level merge by input size: where should the automatic choice switch?"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantik-core-py_b200"))
import quantik_b200 as Q  # noqa: E402
from quantik_b200 import _native as N  # noqa: E402

for log2n in (18, 19, 20, 21, 22, 23, 24, 25):
    pos = Q.random_roots(1 << log2n, min_plies=3, span=6, as_torch=True)
    line = f"2^{log2n}:"
    for mode in (1, 2, 0):
        N.set_option("dedup_mode", mode)
        for _ in range(2):
            u, m = Q.canon_dedup(pos, None, sort_output=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            u, m = Q.canon_dedup(pos, None, sort_output=False)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
        line += f"  mode {mode}: {dt * 1e3:8.3f} ms ({(1 << log2n) / dt:.2e}/s)"
    print(line, int(u.shape[0]), "classes", flush=True)
