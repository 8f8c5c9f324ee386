import sys, time, numpy as np, torch
sys.path.insert(0, "quantik-core-py_b200"); sys.path.insert(0, ".")
import quantik_b200 as Q
from quantik_b200 import batch as B
from oracle import oracle as O
roots = np.ascontiguousarray(B.random_roots(65536, seed=4, min_plies=4, span=8))
won = np.array([1, 2, 4, 8, 0, 0, 0, 0], np.uint16)
for frac in (0.0, 0.25, 0.9):
    r = roots.copy()
    idx = np.random.default_rng(1).random(len(r)) < frac
    r[idx] = won
    rt = torch.from_numpy(r.view(np.int16)).cuda()
    out = {}
    for v in (2, 0):
        Q._native.set_option("playout_variant", v)
        B.playouts(rt, 64)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res = B.playouts(rt, 1024, first_rollout=5)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        out[v] = (res["wins_p0"].cpu().numpy(), res["wins_p1"].cpu().numpy())
        print(f"decided fraction {frac}: variant {v}: {65536 * 1024 / dt:.3e} playouts/s", flush=True)
    assert np.array_equal(out[0][0], out[2][0]) and np.array_equal(out[0][1], out[2][1])
print("tallies equal")
