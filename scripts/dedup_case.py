#!/usr/bin/env python3
"""One 2^24-key canonical dedup per mode (one-level table / two-level merge): ncu launch-list target."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantik-core-py_b200"))
import quantik_b200 as Q  # noqa: E402
from quantik_b200 import _native as N  # noqa: E402

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
pos = Q.random_roots(1 << log2n, min_plies=3, span=6, as_torch=True)
for mode in (1, 2):
    N.set_option("dedup_mode", mode)
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        u, m = Q.canon_dedup(pos, None, sort_output=False)
        torch.cuda.synchronize()
        print(f"mode {mode} rep {rep}: {1e3 * (time.perf_counter() - t0):.3f} ms, {int(u.shape[0])} classes")
N.set_option("dedup_mode", 2)
for streams, budgets in ((1, (96, 112)), (2, (96, 112, 128, 160, 200)), (3, (112,))):
    N.set_option("dedup_streams", streams)
    for mb in budgets:
        N.set_option("dedup_part_mb", mb)
        Q.canon_dedup(pos, None, sort_output=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            u, m = Q.canon_dedup(pos, None, sort_output=False)
        torch.cuda.synchronize()
        print(f"streams {streams}, part table {mb} MiB: {1e3 * (time.perf_counter() - t0) / 5:.3f} ms per call, {int(u.shape[0])} classes")
