#!/usr/bin/env python3
"""Small invocation of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantik-core-py_b200"))
import quantik_b200 as Q  # noqa: E402

e = np.zeros((1, 8), np.uint16)
roots = Q.random_roots(256)
late = Q.random_roots(256, min_plies=8, span=5)
print(Q.playouts(e, 4096, per_playout=True)["wins_p0"], Q.playouts(roots, 32)["wins_p0"][:4], Q.playouts(roots, 16, max_plies=3)["draws"][:4])
for mode in ("warp", "lane"):
    Q._native.set_option("perft_kernel", {"lane": 2, "warp": 3}[mode])
    print(mode, Q.perft(e, 5)["nodes"].tolist(), Q.perft(late, 4)["nodes"].tolist())
Q._native.set_option("perft_kernel", 0)
print(Q.canonical_payloads(roots)[:1], Q.orbit_sizes(roots)[:4], Q.movegen_masks(roots)[0][:2], Q.win_status(roots)[:4])
u, m = Q.canon_dedup(np.repeat(roots, 4, 0))
print(len(u), int(m.sum()))
print(Q.symmetry_table(5), Q.opening_book_summary(3)["total_edges"])
print(Q.guided_playouts(e, 64)["wins_p0"], Q.solve(late)[0][:8], Q.eval_features(roots, np.zeros(256, np.uint8))[:2])
print("sanitizer case done")
