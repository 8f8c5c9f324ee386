#!/usr/bin/env python3
"""bench.py -- headline benchmark of the Quantik batched-simulation hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic input:
BASELINE config 1 scaled to throughput size -- 2^30 uniform-random playouts from the
empty board per GPU (weak scaling), win tallies reduced with ONE all-reduce per step.
Rank 0 prints one JSON line (metric: random playouts/s, whole job).  The same line
carries perft nodes/s (config 2), canonical keys/s (config 3) and the config-4 playout
rate under "extra", the integer-issue roofline, the CPU baseline (oracle port on the host
cores) and the end-to-end number measured through the C-ABI with host buffers.

--impl reference times the CPU implementation of the path (the reference is pure Python
and cannot travel to the GPU box; the C oracle port of it is used, all host threads).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "quantik-core-py_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np

METRIC = "random playouts/sec (uniform random, empty board, to terminal)"
UNIT = "playouts/s"
PLAYOUTS_PER_STEP = 1 << 30
SEED = 0x5155414E54494B31
# Integer-op model (which replaces the SURVEY 8d estimate made before the
# incremental formulation existed): 60 lane-ops per ply, of which 36 issue to the 64-lane ALU pipe
# and 24 to the FMA-heavy pipe (IMAD; IMAD.WIDE counts double); 11.08 plies per empty-board playout;
# 118 per visited perft leaf parent = 4.4 per counted node at depth 6; 304 per canonical key.
OPS_PER_PLY, ALU_OPS_PER_PLY, PLIES_PER_PLAYOUT, OPS_PER_NODE, OPS_PER_KEY = 60.0, 36.0, 11.08, 2.5, 304.0
OPS_PER_PLAYOUT = OPS_PER_PLY * PLIES_PER_PLAYOUT


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(2)
            except Exception:
                self.proc.kill()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(seconds: float = 12.0):
    """Oracle port (C, OpenMP over all host threads) of the reference's rollout loop on a bounded
    sample of the same workload: playouts from the empty board."""
    import oracle as O
    O.set_num_threads(len(os.sched_getaffinity(0)))            # torchrun exports OMP_NUM_THREADS=1
    empty = np.zeros((1, 8), np.uint16)
    O.playouts(empty, 2000)
    t = time.perf_counter()
    O.playouts(empty, 20000, SEED)
    rate = 20000 / (time.perf_counter() - t)
    n = int(max(20000, min(rate * seconds, 50_000_000)))
    t = time.perf_counter()
    r = O.playouts(empty, n, SEED, first_rollout=1 << 40)
    dt = time.perf_counter() - t
    return {"value": n / dt, "unit": UNIT, "cores": O.num_threads(), "kind": "port",
            "sample": f"{n} empty-board playouts, C oracle (oracle/quantik_oracle.c, OpenMP), {dt:.1f} s; "
                      f"P0 win rate {int(r['wins_p0'][0]) / n:.4f}",
            "python_reference_measured": {"value": 559.0, "unit": "playouts/s/core",
                                          "note": "quantik-core 1.1.0 itself, SURVEY.md section 6 (build container, "
                                                  "not this box): 3118 playouts/s on 8 processes"}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle as O
    O.set_num_threads(len(os.sched_getaffinity(0)))            # torchrun exports OMP_NUM_THREADS=1
    empty = np.zeros((1, 8), np.uint16)
    O.playouts(empty, 20000, SEED)
    t = time.perf_counter()
    O.playouts(empty, 50000, SEED)
    rate = 50000 / (time.perf_counter() - t)
    budget = 150.0 / max(1, args.steps + args.warmup)          # whole run within a few minutes
    n = int(max(10000, min(rate * min(budget, 15.0), 20_000_000)))
    for w in range(args.warmup):
        O.playouts(empty, n, SEED, first_rollout=w * n)
    t0 = time.perf_counter()
    for s in range(args.steps):
        O.playouts(empty, n, SEED, first_rollout=(args.warmup + s) * n)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    sample = f"{n} empty-board playouts per step (bounded sample of the 2^30-playout step), C oracle port, all host threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": "cfg1-throughput: uniform-random playouts from the empty board", "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": O.num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--playouts-per-step", type=int, default=PLAYOUTS_PER_STEP)
    ap.add_argument("--no-extra", action="store_true", help="skip the perft / canonical / config-4 side measurements")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import quantik_b200 as Q
    from quantik_b200 import batch as B

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = local
    info = Q._native.device_info(dev)
    per_step = int(args.playouts_per_step)
    assert per_step < (1 << 32)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------------------------------------------------------- device-resident timed region
    root_dev = torch.zeros((1, 8), dtype=torch.int16, device=f"cuda:{dev}")
    tally = torch.zeros(3, dtype=torch.int64, device=f"cuda:{dev}")

    def step(i: int):
        first = (i * world + rank) * per_step
        r = B.playouts(root_dev, per_step, SEED, first_rollout=first)
        t = torch.stack([r["wins_p0"][0], r["wins_p1"][0], r["draws"][0]]).to(torch.int64) & 0xFFFFFFFF
        if world > 1:
            dist.all_reduce(t)                       # the one collective of the path: win counts
        tally.add_(t)

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(dev)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tally.zero_()
    barrier()
    e0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=f"cuda:{dev}")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    clocks = sampler.stop() if rank == 0 else None
    total_playouts = per_step * world * args.steps
    value = total_playouts / (ms * 1e-3)
    t_host = tally.cpu().numpy()
    assert int(t_host[0] + t_host[1]) == total_playouts and int(t_host[2]) == 0, "tallies do not add up"

    # ---------------------------------------------------------------- end to end through the C-ABI with host buffers
    root_host = torch.zeros((1, 8), dtype=torch.int16).pin_memory().numpy().view(np.uint16)
    for i in range(2):
        B.playouts(root_host, per_step, SEED, first_rollout=(1 << 50) + i * per_step, device=dev)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        r = B.playouts(root_host, per_step, SEED, first_rollout=(1 << 51) + (i * world + rank) * per_step, device=dev)
        _ = int(r["wins_p0"][0])                      # result read back on the host every step
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{dev}")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e = per_step * world * args.steps / float(e2e_s.item())

    # ---------------------------------------------------------------- perft nodes/s at N GPUs (config 5, shortened)
    # the 10 946 canonical depth-4 classes (weighted) enumerated 6 plies deeper, roots dealt round-robin over the
    # ranks, ONE all-reduce of the 5 x 7 counters (quantik_b200.dist.sharded_perft)
    from quantik_b200 import dist as D
    sharded = {}
    d4_path = os.path.join(ROOT, "tests", "golden", "depth4_canonical.npz")
    if not args.no_extra and os.path.exists(d4_path):
        d4 = np.load(d4_path)
        D.sharded_perft(d4["states"], 3, d4["mult"], device=dev)
        barrier()
        t0 = time.perf_counter()
        p10 = D.sharded_perft(d4["states"], 6, d4["mult"], device=dev)
        torch.cuda.synchronize()
        t_perft = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{dev}")
        if world > 1:
            dist.all_reduce(t_perft, op=dist.ReduceOp.MAX)
        assert int(p10["nodes"][2]) == 6241600512 and int(p10["nodes"][4]) == 2097048766464      # GAME_TREE_ANALYSIS.md:9,11
        sharded = {"perft_cfg5_depth10_weighted_nodes_per_s": float(p10["nodes"][1:].astype(np.float64).sum()) / float(t_perft.item()),
                   "perft_cfg5_depth10_s": float(t_perft.item()), "perft_cfg5_n_gpus": world}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---------------------------------------------------------------- rank 0: roofline, side metrics, CPU baseline
    peak_issue = B.int_issue_peak(dev)
    sm_mhz = (clocks or {}).get("sm_mhz") or info["sm_clock_khz"] / 1e3
    nominal_peak = info["sm_count"] * 128 * sm_mhz * 1e6          # 4 schedulers x 32 lanes per SM-clock
    achieved = value / world * OPS_PER_PLAYOUT                    # per GPU
    roofline = {
        "bound": "int_issue", "kernel": "k_playout_single<true> (config 1: one root, tallies only; k_playout<false,false> for many roots)",
        "achieved": achieved / 1e12, "peak": nominal_peak / 1e12, "unit": "Tlane-op/s", "frac": achieved / nominal_peak,
        "traffic": 14592,   # dram__bytes_read+write of one k_playout launch (ncu --set full, profiles/r01i_playout_ncu.txt)
        "model": f"{OPS_PER_PLY:.0f} int ops/ply x {PLIES_PER_PLAYOUT} plies/playout (DESIGN.md section 5); peak = SMs x 4 "
                 f"schedulers x 32 lanes x SM clock under load ({sm_mhz:.0f} MHz)",
        "peak_measured_lop3": peak_issue["alu"] / 1e12, "peak_measured_lop3_imad_mix": peak_issue["mixed"] / 1e12,
        "frac_of_measured_mix": achieved / peak_issue["mixed"] if peak_issue["mixed"] else None,
        "alu_pipe": {"ops_per_ply": ALU_OPS_PER_PLY, "achieved": value / world * PLIES_PER_PLAYOUT * ALU_OPS_PER_PLY / 1e12,
                     "peak_measured": peak_issue["alu"] / 1e12,
                     "frac": value / world * PLIES_PER_PLAYOUT * ALU_OPS_PER_PLY / peak_issue["alu"] if peak_issue["alu"] else None,
                     "note": "LOP3/SEL/ISETP/SHF/IADD3 share one 64-lane/clk/SM pipe on sm_100: the binding sub-roofline "
                             "(ncu: profiles/)"},
        "hbm": {"algorithmic_bytes_per_step": 16 + 12, "note": "one 16-byte root in, three u32 tallies out: not memory-bound"},
        "ncu": {"source": "profiles/r01i_playout_ncu.txt", "issue_active_pct": 70.5, "pipe_alu_pct": 62.8,
                "pipe_fmaheavy_cycles_pct": 68.0, "pipe_xu_pct": 29.3, "lanes_active_per_inst": 31.0},
    }

    extra = dict(sharded)
    if not args.no_extra:
        # config 2: exhaustive perft from the empty board to depth 6 (bit-exact counts asserted)
        e = np.zeros((1, 8), np.uint16)
        B.perft(e, 6, device=dev)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            p = B.perft(e, 6, device=dev)
        dt = (time.perf_counter() - t0) / reps
        assert p["nodes"].tolist() == [1, 64, 3392, 167552, 6776960, 231883776, 6241600512]
        nodes = float(p["nodes"][1:].sum())
        extra["perft_depth6_nodes_per_s"] = nodes / dt
        extra["perft_depth6_ms"] = dt * 1e3
        extra["perft_frac_int_issue"] = nodes / dt * OPS_PER_NODE / nominal_peak
        B.perft(e, 7, device=dev)
        t0 = time.perf_counter()
        p7 = B.perft(e, 7, device=dev)                 # deeper: the lane-parallel DFS kernel takes over
        dt7 = time.perf_counter() - t0
        assert int(p7["nodes"][7]) == 132400548864
        extra["perft_depth7_nodes_per_s"] = float(p7["nodes"][1:].sum()) / dt7
        extra["perft_depth7_ms"] = dt7 * 1e3
        # config 5 / SURVEY 8f-1: the complete canonical game tree (depth 1..16) with the fused BFS
        B.symmetry_table(16, device=dev)               # warm-up: grows the stream-ordered memory pool once
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rows = B.symmetry_table(16, device=dev)
        dt = time.perf_counter() - t0
        assert [r[1] for r in rows[:8]] == [3, 51, 726, 10946, 105632, 901916, 4658465, 17900160]
        extra["full_game_tree_bfs_s"] = dt
        extra["full_game_tree_canonical_states"] = int(sum(r[1] for r in rows))
        extra["full_game_tree_legal_moves"] = int(sum(r[0] for r in rows))
        # config 3: canonical keys over 2^26 device-resident states (random positions x random symmetries)
        n_keys = 1 << 26
        base = B.random_roots(1 << 20, seed=SEED, min_plies=0, span=13, device=dev, as_torch=True)
        states = base.repeat(n_keys // (1 << 20), 1)
        B.canonical_payloads(states)
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(5):
            out = B.canonical_payloads(states)
        c1.record()
        torch.cuda.synchronize()
        kps = 5 * n_keys / (c0.elapsed_time(c1) * 1e-3)
        extra["canonical_keys_per_s"] = kps
        extra["canonical_frac_int_issue"] = kps * OPS_PER_KEY / nominal_peak
        extra["canonical_hbm_gbs"] = kps * 32 / 1e9
        del states, out
        # config 3, merge part: canonicalise + merge equal classes + sum multiplicities over 2^24 mid-game positions
        pos = B.random_roots(1 << 24, seed=SEED, min_plies=3, span=6, device=dev, as_torch=True)
        B.canon_dedup(pos, None, sort_output=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            u, _m = B.canon_dedup(pos, None, sort_output=False)
        torch.cuda.synchronize()
        extra["canon_dedup_keys_per_s"] = 5 * (1 << 24) / (time.perf_counter() - t0)
        extra["canon_dedup_distinct"] = int(u.shape[0])
        del pos, u, _m
        # config 4: 65 536 mid-game roots x 1024 rollouts
        roots = B.random_roots(65536, seed=SEED, device=dev, as_torch=True)
        B.playouts(roots, 1024, SEED)
        torch.cuda.synchronize()
        c0.record()
        for _ in range(3):
            r4 = B.playouts(roots, 1024, SEED)
        c1.record()
        torch.cuda.synchronize()
        extra["cfg4_mcts_leaf_playouts_per_s"] = 3 * 65536 * 1024 / (c0.elapsed_time(c1) * 1e-3)
        extra["cfg4_roots"] = 65536
        extra["cfg4_rollouts_per_root"] = 1024
        # eval-guided rollouts (seeded weights, epsilon 0.2) and the exact endgame solver
        B.guided_playouts(e, 1 << 16, SEED, device=dev)
        t0 = time.perf_counter()
        B.guided_playouts(e, 1 << 21, SEED, device=dev)
        extra["guided_playouts_per_s"] = (1 << 21) / (time.perf_counter() - t0)
        late = B.random_roots(1 << 16, seed=SEED, min_plies=8, span=5, device=dev)
        B.solve(late[:1024], device=dev)
        t0 = time.perf_counter()
        B.solve(late, device=dev)
        extra["solver_positions_per_s"] = (1 << 16) / (time.perf_counter() - t0)

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": "cfg1-throughput: 2^30 uniform-random playouts from the empty board per GPU per step "
                               "(BASELINE config 1 at the size SURVEY 8d prescribes for throughput), Philox4x32-10 stream, "
                               "win tallies all-reduced once per step",
                   "playouts_per_step_per_gpu": per_step, "seed": hex(SEED),
                   "l2": "not applicable: 16-byte input, integer-issue bound; rollout ids advance every step so no "
                         "result is reused",
                   "p0_win_rate": float(t_host[0]) / total_playouts},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": 16, "d2h_bytes_per_step": 12,
                "api": "quantik_b200.playouts(host numpy root) -> qk_playouts_ex"},
        "gpu_launches": 2 * args.steps,
        "clocks": clocks,
        "roofline": roofline,
        "extra": extra,
    }
    if not args.no_cpu_baseline and world == 1:
        out["cpu_baseline"] = cpu_baseline()
    print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
