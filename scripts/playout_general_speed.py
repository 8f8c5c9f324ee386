import sys, time, torch
sys.path.insert(0, "quantik-core-py_b200")
import quantik_b200 as Q
from quantik_b200 import batch as B
roots = B.random_roots(65536, seed=4, min_plies=4, span=8, as_torch=True)
for label, kw in (("cut-off 16", dict(max_plies=16)), ("cut-off 6", dict(max_plies=6)), ("tallies", dict())):
    for variant in (2, 0):
        Q._native.set_option("playout_variant", variant)
        B.playouts(roots, 64, **kw)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(3):
            r = B.playouts(roots, 1024, first_rollout=i * 1024, **kw)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        print(f"{label:12s} variant {variant}: {65536 * 1024 / dt:.4e} playouts/s draws {float(r['draws'].sum()) / (65536 * 1024):.4f}", flush=True)
small = B.random_roots(1 << 20, seed=5, min_plies=4, span=8, as_torch=True)
for variant in (2, 0, 5):
    Q._native.set_option("playout_variant", variant)
    for rollouts in (1, 4, 16):
        B.playouts(small, rollouts)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(3):
            r = B.playouts(small, rollouts, first_rollout=i * 16)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        print(f"2^20 roots x {rollouts:2d} variant {variant}: {(1 << 20) * rollouts / dt:.4e} playouts/s", flush=True)
