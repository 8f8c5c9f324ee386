#!/usr/bin/env python3
"""When does every warp of the playout kernel start and finish?  Needs the experiment build
(nvcc ... -DQK_PLAYOUT_TIMELINE -o quantik_b200/_lib/libquantik_b200_timeline.so); run with
QUANTIK_B200_LIB=<that file> python scripts/playout_timeline.py <variant>."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantik-core-py_b200"))
import quantik_b200 as Q  # noqa: E402
from quantik_b200 import _native as N  # noqa: E402

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 5
n = 1 << 28
N.set_option("playout_variant", variant)
root = torch.zeros((1, 8), dtype=torch.int16, device="cuda:0")
for i in range(3):
    Q.playouts(root, n, first_rollout=i * n)
torch.cuda.synchronize()
lib = ctypes.CDLL(N.LIB_PATH)
buf = np.zeros(4 * 16384, np.uint64)
rc = lib.qk_debug_timeline(buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(buf.size))
t = buf.reshape(-1, 4)
t = t[t[:, 1] > 0]
t0 = t[:, 0].min()
beg, end, sm = (t[:, 0] - t0) / 1e6, (t[:, 1] - t0) / 1e6, t[:, 2]
print("rc", rc, "warps", len(t), "SMs", len(np.unique(sm)))
print("begin ms: min %.3f median %.3f max %.3f" % (beg.min(), np.median(beg), beg.max()))
print("end   ms: min %.3f p10 %.3f median %.3f p90 %.3f max %.3f" % (end.min(), np.percentile(end, 10), np.median(end), np.percentile(end, 90), end.max()))
print("mean warp busy fraction of the kernel: %.3f" % ((end - beg).mean() / end.max()))
per_sm = {}
for s in np.unique(sm):
    m = sm == s
    per_sm[int(s)] = (int(m.sum()), float(end[m].max()))
cnt = np.array([v[0] for v in per_sm.values()])
print("warps per SM: min %d max %d;" % (cnt.min(), cnt.max()), "SM finish ms: min %.3f max %.3f" % (min(v[1] for v in per_sm.values()), max(v[1] for v in per_sm.values())))
hist, edges = np.histogram(end, bins=10)
print("end histogram:", list(zip(np.round(edges[:-1], 2), hist)))
