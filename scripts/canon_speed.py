#!/usr/bin/env python3
"""canonical keys/s over 2^26 device-resident states + the merge paths that canonicalise (dedup, complete BFS)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantik-core-py_b200"))
import quantik_b200 as Q  # noqa: E402
from quantik_b200 import batch as B  # noqa: E402

for a in sys.argv[1:]:
    if a.startswith("waves="):
        Q._native.set_option("grid_waves", int(a[6:]))
base = B.random_roots(1 << 20, min_plies=0, span=13, as_torch=True)
states = base.repeat(64, 1)
for cs in (False, True):
    B.canonical_payloads(states, color_swap=cs)
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(5):
        B.canonical_payloads(states, color_swap=cs)
    c1.record()
    torch.cuda.synchronize()
    print(f"canonical ({'384' if cs else '192'} images): {5 * (1 << 26) / (c0.elapsed_time(c1) * 1e-3):.4e} keys/s", flush=True)
pos = B.random_roots(1 << 24, min_plies=3, span=6, as_torch=True)
B.canon_dedup(pos, None, sort_output=False)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    B.canon_dedup(pos, None, sort_output=False)
torch.cuda.synchronize()
print(f"dedup 2^24: {5 * (1 << 24) / (time.perf_counter() - t0):.4e} keys/s", flush=True)
Q.symmetry_table(16)
torch.cuda.synchronize()
t0 = time.perf_counter()
Q.symmetry_table(16)
torch.cuda.synchronize()
print(f"complete game tree: {time.perf_counter() - t0:.4f} s", flush=True)
