import sys, time, torch
sys.path.insert(0, "quantik-core-py_b200")
import quantik_b200 as Q
from quantik_b200 import batch as B
roots = B.random_roots(65536, seed=4, min_plies=4, span=8, as_torch=True)
for grab in (16, 32, 64, 128, 256):
    Q._native.set_option("playout_grab", grab)
    B.playouts(roots, 64)
    ts = []
    for i in range(7):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = B.playouts(roots, 1024, first_rollout=i * 1024)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print(f"config 4 grab {grab:3d}: {65536 * 1024 / sorted(ts)[3]:.4e} playouts/s", flush=True)
