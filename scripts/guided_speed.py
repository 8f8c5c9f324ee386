import sys, time, numpy as np
sys.path.insert(0, "quantik-core-py_b200")
import quantik_b200 as Q
from quantik_b200 import batch as B
e = np.zeros((1, 8), np.uint16)
Q.guided_playouts(e, 1 << 16)
for waves in (1, 2, 4, 8, 32):
    Q._native.set_option("grid_waves", waves)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); r = Q.guided_playouts(e, 1 << 24); ts.append(time.perf_counter() - t0)
    print("grid_waves", waves, f"{(1 << 24) / min(ts):.4e}")
Q._native.set_option("grid_waves", 0)
for scoring in (1, 0):
    Q._native.set_option("guided_scoring", scoring)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); r = Q.guided_playouts(e, 1 << 24); ts.append(time.perf_counter() - t0)
    print("guided_scoring", scoring, f"{(1 << 24) / min(ts):.4e}")
for n in (1 << 22, 1 << 24):
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); r = Q.guided_playouts(e, n); ts.append(time.perf_counter() - t0)
    print("empty board", n, f"{n / min(ts):.4e}", int(r["wins_p0"][0]), int(r["wins_p1"][0]))
roots = B.random_roots(4096, seed=4, min_plies=4, span=8)
Q.guided_playouts(roots, 64)
for scoring in (1, 0):
    Q._native.set_option("guided_scoring", scoring)
    t0 = time.perf_counter(); r = Q.guided_playouts(roots, 4096); dt = time.perf_counter() - t0
    print("mid-game roots, guided_scoring", scoring, f"{4096 * 4096 / dt:.4e}")
