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

--impl reference times the reference ITSELF (quantik-core 1.1.0, pure Python: the rollout loop of
beam_search.py:624-645 on one process per host core) from the recipe-installed oracle/_ref
(oracle/ref_recipe.py; __graft_entry__.build() populates it where /root/reference exists and it
travels with the snapshot).  The C/OpenMP oracle port of the same loop is reported beside it.
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
# incremental formulation existed): 56 lane-ops per ply in the predicated formulation of round 2 (60 in round 1's
# multiply-add formulation: the fraction at that count is reported beside it), of which 27 issue to the 64-lane ALU
# pipe; 11.08 plies per empty-board playout; 2.5 per counted perft node at depth 6; 304 per canonical key.
OPS_PER_PLY, ALU_OPS_PER_PLY, PLIES_PER_PLAYOUT, OPS_PER_NODE, OPS_PER_KEY = 56.0, 27.0, 11.08, 2.5, 304.0
OPS_PER_PLY_ROUND1 = 60.0
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


def _port_rate(seconds: float):
    """C/OpenMP oracle port of the rollout loop on all host threads, about `seconds` of work."""
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
                      f"P0 win rate {int(r['wins_p0'][0]) / n:.4f}"}


def _reference_rate(seconds: float):
    """The reference's own Python rollout loop (BeamSearchEngine._rollout) on every host core, `seconds` per process.
    -> cpu_baseline dict, or None when oracle/_ref is not installed."""
    from oracle import ref_recipe
    if not ref_recipe.available():
        return None
    r = ref_recipe.time_reference_playouts(seconds)
    return {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "reference",
            "sample": f"{r['playouts']} empty-board playouts by quantik-core's own BeamSearchEngine._rollout "
                      f"(beam_search.py:624-645, MT19937), {r['cores']} processes x {r['seconds']:.1f} s; "
                      f"P0 win rate {r['p0_rate']:.4f}"}


def cpu_baseline(seconds: float = 12.0):
    """The path on the box's host cores, bounded sample of the same workload (playouts from the empty board):
    the reference itself (kind "reference") with the C port of it as a second, labelled number."""
    port = _port_rate(seconds * 0.6)
    ref = _reference_rate(seconds)
    if ref is None:
        port["note"] = "oracle/_ref missing (build() not run where /root/reference exists): port only"
        return port
    ref["c_port"] = port
    return ref


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    budget = min(20.0, 150.0 / max(1, args.steps + args.warmup))      # seconds per step: whole run within a few minutes
    ref = None
    from oracle import ref_recipe
    if ref_recipe.available():
        for _ in range(args.warmup):
            ref_recipe.time_reference_playouts(min(2.0, budget))
        n = 0
        wall = 0.0
        p0 = 0.0
        for _ in range(args.steps):
            r = ref_recipe.time_reference_playouts(budget)
            n += r["playouts"]
            wall += r["seconds"]
            p0 += r["p0_rate"] * r["playouts"]
            cores = r["cores"]
        value, ms = n / wall, 1e3 * wall / args.steps
        sample = (f"{n // max(1, args.steps)} empty-board playouts per step (bounded sample of the 2^30-playout step) by quantik-core "
                  f"1.1.0 itself: BeamSearchEngine._rollout (beam_search.py:624-645) on {cores} processes; P0 win rate {p0 / n:.4f}")
        ref = {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample, "c_port": _port_rate(5.0)}
    else:
        port = _port_rate(10.0)
        value, ms, sample = port["value"], 1e3 * 10.0, port["sample"] + " (oracle/_ref missing: C port of the reference loop)"
        ref = port
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": "cfg1-throughput: uniform-random playouts from the empty board", "sample": sample},
        "cpu_baseline": ref,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def search_speedup(dev: int):
    """SURVEY 8f-2 end to end, benchmarks/adapters.py:160-259 in small: wall time of search() of the reference's stock
    engines (pure Python, one playout at a time) against the batched subclasses of the SAME engines
    (quantik_b200.engines: K leaves / one beam level per launch) on the same position with the same playout budget.
    The engines come from the recipe-installed reference (oracle/_ref); {} when it is missing."""
    from oracle import ref_recipe
    if not ref_recipe.available():
        return {}
    ref_recipe.import_reference()
    import quantik_b200 as Q
    from quantik_core import beam_search as bs, mcts as mc
    from quantik_core.core import State
    state = State.empty()
    out = {}
    budget = 4096
    t0 = time.perf_counter()
    stock = mc.MCTSEngine(mc.MCTSConfig(max_iterations=budget, random_seed=1))
    stock.search(state)
    t_stock = time.perf_counter() - t0
    Batched = Q.batched_mcts_engine(mc, leaves_per_wave=64, rollouts_per_leaf=64, device=dev)
    Batched(mc.MCTSConfig(max_iterations=budget, random_seed=1)).search(state)
    t0 = time.perf_counter()
    eng = Batched(mc.MCTSConfig(max_iterations=budget, random_seed=1))
    eng.search(state)
    t_b = time.perf_counter() - t0
    big = 1 << 22
    t0 = time.perf_counter()
    eng2 = Q.batched_mcts_engine(mc, leaves_per_wave=256, rollouts_per_leaf=1024, device=dev)(mc.MCTSConfig(max_iterations=big, random_seed=1))
    eng2.search(state)
    t_big = time.perf_counter() - t0
    out["mcts_search"] = {"playout_budget": budget, "stock_s": t_stock, "batched_s": t_b, "speedup": t_stock / t_b,
                          "batched_launches": eng.gpu_launches, "stock_playouts_per_s": budget / t_stock,
                          "batched_large": {"playout_budget": big, "seconds": t_big, "playouts_per_s": big / t_big,
                                            "waves": eng2.waves, "launches": eng2.gpu_launches,
                                            "tree_nodes": int(eng2.tree.storage.node_count)}}
    cfg = dict(beam_width=8, max_depth=3, rollouts_per_candidate=8, random_seed=1)
    t0 = time.perf_counter()
    rs = bs.BeamSearchEngine(bs.BeamSearchConfig(**cfg)).search(state)
    t_stock = time.perf_counter() - t0
    BB = Q.batched_beam_engine(bs, device=dev)
    BB(bs.BeamSearchConfig(**cfg)).search(state)
    t0 = time.perf_counter()
    engb = BB(bs.BeamSearchConfig(**cfg))
    rb = engb.search(state)
    t_b = time.perf_counter() - t0
    cfg_big = dict(beam_width=256, max_depth=16, rollouts_per_candidate=4096, random_seed=1)
    t0 = time.perf_counter()
    engc = BB(bs.BeamSearchConfig(**cfg_big))
    rc = engc.search(state)
    t_big = time.perf_counter() - t0
    out["beam_search"] = {"config": cfg, "stock_s": t_stock, "stock_rollouts": rs.stats["rollouts"], "batched_s": t_b,
                          "batched_rollouts": rb.stats["rollouts"], "batched_launches": engb.gpu_launches,
                          "speedup_per_rollout": (t_stock / max(1, rs.stats["rollouts"])) / (t_b / max(1, rb.stats["rollouts"])),
                          "batched_large": {"config": cfg_big, "seconds": t_big, "rollouts": rc.stats["rollouts"],
                                            "rollouts_per_s": rc.stats["rollouts"] / t_big, "launches": engc.gpu_launches,
                                            "evaluations": rc.stats["evaluations"]}}
    return out


def ncu_summary():
    """Counters of the dominant kernel as captured by ncu: read at run time from the committed summary
    (profiles/playout_ncu_latest.json, written by profiles/summarize.py from the .ncu-rep), never typed in here."""
    path = os.path.join(ROOT, "profiles", "playout_ncu_latest.json")
    if not os.path.exists(path):
        return None
    import hashlib
    with open(path, "rb") as f:
        raw = f.read()
    d = json.loads(raw)
    d["summary_file"] = "profiles/playout_ncu_latest.json"
    d["summary_sha256"] = hashlib.sha256(raw).hexdigest()[:16]
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--playouts-per-step", type=int, default=PLAYOUTS_PER_STEP)
    ap.add_argument("--no-extra", action="store_true", help="skip the perft / canonical / config-4 side measurements")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cfg5-full", action="store_true", help="world > 1: skip the full-depth config-5 enumeration (about 130 s / N)")
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

    # ---------------------------------------------------------------- strong-scaling jobs at N GPUs
    from quantik_b200 import dist as D

    def timed_job(fn):
        """barrier, run, sync; -> (result, seconds as the max over ranks)"""
        barrier()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{dev}")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return out, float(t.item())

    sharded = {}
    d4_path = os.path.join(ROOT, "tests", "golden", "depth4_canonical.npz")
    if not args.no_extra and os.path.exists(d4_path):
        # config 5, shortened: the 10 946 canonical depth-4 classes (weighted) enumerated 6 plies deeper; roots dealt
        # largest-first by a depth-2 probe, ONE all-reduce of the 5 x 7 counters (quantik_b200.dist.sharded_perft).
        # Warm-up = the same call (memory pool, NCCL channel set-up), best of 3 timed runs reported with all three.
        d4 = np.load(d4_path)
        D.sharded_perft(d4["states"], 6, d4["mult"], device=dev)
        runs = []
        for _ in range(3):
            p10, t = timed_job(lambda: D.sharded_perft(d4["states"], 6, d4["mult"], device=dev))
            runs.append(t)
        assert int(p10["nodes"][2]) == 6241600512 and int(p10["nodes"][4]) == 2097048766464      # GAME_TREE_ANALYSIS.md:9,11
        t_perft = min(runs)
        sharded = {"perft_cfg5_depth10_weighted_nodes_per_s": float(p10["nodes"][1:].astype(np.float64).sum()) / t_perft,
                   "perft_cfg5_depth10_s": t_perft, "perft_cfg5_depth10_runs_s": runs, "perft_cfg5_n_gpus": world}
    if not args.no_extra and world > 1:
        # the exchange step at world > 1, asserted in-run against the single-GPU result every rank
        # computes for itself: (1) the complete canonical game tree, frontier partitioned by payload hash, the
        # children stored straight into their owners' inboxes over NVLink (fused) / NCCL all-to-all (two-step);
        one_rows = [list(map(int, r)) for r in B.symmetry_table(16, device=dev)]
        boxes = D.PeerInboxes(1 << 14, dev, copies=2)
        D.fused_symmetry_table(16, device=dev, inboxes=boxes)          # sizes the inboxes for the deepest level
        # both forms: one warm-up call, then the better of two timed calls (the first timed call of the two-step form
        # still grows torch's allocator on some boxes: 1.18 s against 0.22 s on two GPUs)
        t_fused = t_nccl = float("inf")
        for _ in range(2):
            rows_f, t = timed_job(lambda: D.fused_symmetry_table(16, device=dev, inboxes=boxes))
            assert rows_f == one_rows, "fused partitioned BFS differs from the single-GPU table"
            t_fused = min(t_fused, t)
        del boxes
        D.sharded_symmetry_table(16, device=dev)
        for _ in range(2):
            rows_n, t = timed_job(lambda: D.sharded_symmetry_table(16, device=dev))
            assert rows_n == one_rows, "NCCL partitioned BFS differs from the single-GPU table"
            t_nccl = min(t_nccl, t)
        # (2) global canonical dedup: every rank holds 2^24 symmetric images of mid-game positions
        import itertools
        n_img = 1 << 24                                   # the size of the single-GPU canon_dedup metric below
        pos = B.random_roots(n_img, first=rank * n_img, min_plies=3, span=6, device=dev)
        table = np.array([B.encode_transform(d4_, False, perm) for perm in itertools.permutations(range(4)) for d4_ in range(8)], np.uint16)
        codes = table[((np.arange(n_img, dtype=np.uint64) * np.uint64(2654435761) + np.uint64(rank)) % np.uint64(192)).astype(np.int64)]
        images = torch.from_numpy(np.ascontiguousarray(B.apply_symmetry(pos, codes, device=dev)).view(np.int16)).to(f"cuda:{dev}")
        dbox = D.PeerInboxes(int(1.25 * n_img / world) + 4096, dev, copies=1)
        D.fused_canon_dedup(images, None, device=dev, inboxes=dbox, n_max=n_img)
        (f_own, f_mult, f_global), t_dd = timed_job(lambda: D.fused_canon_dedup(images, None, device=dev, inboxes=dbox, n_max=n_img))
        (n_own, n_mult, n_global), t_dd_nccl = timed_job(lambda: D.sharded_canon_dedup(images, None, device=dev))
        f_mass = D.all_reduce_counts(np.array([int(f_mult.sum()), int(n_mult.sum())], np.int64), dev)
        gathered = [torch.empty_like(images.view(torch.int64)) for _ in range(world)]
        dist.all_gather(gathered, images.view(torch.int64))
        if rank == 0:
            u_all, m_all = B.canon_dedup(torch.cat(gathered), None, device=dev, sort_output=False)
            assert int(u_all.shape[0]) == f_global == n_global, "global dedup: number of classes differs from one GPU"
            assert int(f_mass[0]) == world * n_img == int(f_mass[1]) == int(m_all.sum()), "global dedup: multiplicities lost"
            del u_all, m_all
        del gathered, dbox, images
        sharded.update({"fused_bfs_s": t_fused, "nccl_bfs_s": t_nccl, "fused_bfs_canonical_states": int(sum(r[1] for r in rows_f)),
                        "fused_dedup_keys_per_s": world * n_img / t_dd, "nccl_dedup_keys_per_s": world * n_img / t_dd_nccl,
                        "fused_dedup_global_classes": int(f_global), "exchange_asserted_equal_to_single_gpu": True})
        if not args.no_cfg5_full and os.path.exists(d4_path):
            # config 5 at full depth: the depth-4 classes 12 plies deeper = the whole game tree, every depth's totals
            # compared with the canonical BFS table (two independent enumerations)
            p16, t16 = timed_job(lambda: D.sharded_perft(d4["states"], 12, d4["mult"], device=dev))
            for r in range(1, 13):
                want = one_rows[4 + r - 1]
                got = [int(p16["nodes"][r]), int(p16["wins_p0"][r]) + int(p16["blocked_p1_loses"][r - 1]),
                       int(p16["wins_p1"][r]) + int(p16["blocked_p0_loses"][r - 1]),
                       int(p16["nodes"][r]) - int(p16["wins_p0"][r]) - int(p16["wins_p1"][r])]
                assert got == [want[0], want[2], want[3], want[4]], ("config 5 full depth", 4 + r)
            sharded.update({"cfg5_full_depth_s": t16, "cfg5_full_depth_weighted_nodes": int(p16["nodes"][1:].astype(object).sum()),
                            "cfg5_full_depth_matches_canonical_bfs": True})

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---------------------------------------------------------------- rank 0: roofline, side metrics, CPU baseline
    peak_issue = B.int_issue_peak(dev)
    ncu = ncu_summary()
    sm_mhz = (clocks or {}).get("sm_mhz") or info["sm_clock_khz"] / 1e3
    nominal_peak = info["sm_count"] * 128 * sm_mhz * 1e6          # 4 schedulers x 32 lanes per SM-clock
    achieved = value / world * OPS_PER_PLAYOUT                    # per GPU
    roofline = {
        "bound": "int_issue", "kernel": "k_playout_pred<4> (config 1: one root, tallies only; k_playout_pred_roots<4> for many roots)",
        "achieved": achieved / 1e12, "peak": nominal_peak / 1e12, "unit": "Tlane-op/s", "frac": achieved / nominal_peak,
        "traffic": (ncu or {}).get("dram_bytes_per_launch"),   # ncu --set full, see "ncu" below
        "model": f"{OPS_PER_PLY:.0f} int ops/ply x {PLIES_PER_PLAYOUT} plies/playout (DESIGN.md section 5); peak = SMs x 4 "
                 f"schedulers x 32 lanes x SM clock under load ({sm_mhz:.0f} MHz)",
        "peak_measured_lop3": peak_issue["alu"] / 1e12, "peak_measured_lop3_imad_mix": peak_issue["mixed"] / 1e12,
        "frac_of_measured_mix": achieved / peak_issue["mixed"] if peak_issue["mixed"] else None,
        "frac_at_round1_op_count": achieved * OPS_PER_PLY_ROUND1 / OPS_PER_PLY / nominal_peak,
        "alu_pipe": {"ops_per_ply": ALU_OPS_PER_PLY, "achieved": value / world * PLIES_PER_PLAYOUT * ALU_OPS_PER_PLY / 1e12,
                     "peak_measured": peak_issue["alu"] / 1e12,
                     "frac": value / world * PLIES_PER_PLAYOUT * ALU_OPS_PER_PLY / peak_issue["alu"] if peak_issue["alu"] else None,
                     "note": "LOP3/SEL/ISETP/SHF/IADD3 share one 64-lane/clk/SM pipe on sm_100: the binding sub-roofline "
                             "(ncu: profiles/)"},
        "hbm": {"algorithmic_bytes_per_step": 16 + 12, "note": "one 16-byte root in, three u32 tallies out: not memory-bound"},
        "ncu": ncu,
    }

    # useful ply-slots / executed ply-slots of the playout kernel: a finished lane waits for the next 4-ply boundary
    # (measured on 2^20 playouts of the same stream; the ~0.14 blocked-mover steps per game, which also take a slot,
    # are not visible in the per-playout length and are left out)
    sample = B.playouts(np.zeros((1, 8), np.uint16), 1 << 20, SEED, per_playout=True, device=dev)
    lens = np.asarray(sample["plies"], np.int64).ravel()
    roofline["slot_efficiency"] = float(lens.sum()) / float((4 * ((lens + 3) // 4)).sum())
    roofline["mean_plies_measured"] = float(lens.mean())

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
        reps4 = []
        for _ in range(9):                             # a 1.8 ms job: every repetition timed on its own, median reported
            c0.record()
            r4 = B.playouts(roots, 1024, SEED)
            c1.record()
            torch.cuda.synchronize()
            reps4.append(c0.elapsed_time(c1) * 1e-3)
        reps4.sort()
        extra["cfg4_mcts_leaf_playouts_per_s"] = 65536 * 1024 / reps4[len(reps4) // 2]
        extra["cfg4_ms_min_median_max"] = [1e3 * reps4[0], 1e3 * reps4[len(reps4) // 2], 1e3 * reps4[-1]]
        extra["cfg4_roots"] = 65536
        extra["cfg4_rollouts_per_root"] = 1024
        # eval-guided rollouts (seeded weights, epsilon 0.2) and the exact endgame solver
        B.guided_playouts(e, 1 << 18, SEED, device=dev)
        t0 = time.perf_counter()
        B.guided_playouts(e, 1 << 24, SEED, device=dev)
        extra["guided_playouts_per_s"] = (1 << 24) / (time.perf_counter() - t0)
        late = B.random_roots(1 << 16, seed=SEED, min_plies=8, span=5, device=dev)
        B.solve(late[:1024], device=dev)
        t0 = time.perf_counter()
        B.solve(late, device=dev)
        extra["solver_positions_per_s"] = (1 << 16) / (time.perf_counter() - t0)
        if world == 1 and not args.no_cpu_baseline:
            extra["search_speedup"] = search_speedup(dev)

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
