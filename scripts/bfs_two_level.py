#!/usr/bin/env python3
"""Complete canonical game tree: one-level merge table against the routed (two-level) form, part-table sizes, streams."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantik-core-py_b200"))
import quantik_b200 as Q  # noqa: E402
from quantik_b200 import _native as N  # noqa: E402

N.init(0)
ref = None


def run(label):
    global ref
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rows = Q.symmetry_table(16)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    if ref is None:
        ref = rows
    assert rows == ref, label
    print(f"{label}: bfs(16) {best:.4f} s, states {sum(r[1] for r in rows)}", flush=True)


N.set_option("dedup_mode", 1)
run("one-level table")
N.set_option("dedup_mode", 0)
for streams in (1, 2, 3):
    N.set_option("dedup_streams", streams)
    for mb in (48, 64, 96, 112, 160):
        N.set_option("dedup_part_mb", mb)
        run(f"two-level (auto), {streams} stream(s), part table {mb} MiB")
