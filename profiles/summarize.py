#!/usr/bin/env python3
"""Turn an .ncu-rep (ncu --set full) into the short text summary committed under profiles/.
Usage: python profiles/summarize.py gpurun_out/prof.ncu-rep > profiles/<name>.txt
       python profiles/summarize.py gpurun_out/prof.ncu-rep --json profiles/playout_ncu_latest.json [units-per-launch]
(the second form writes the machine-readable summary bench.py loads at run time for its roofline block)"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.avg.per_cycle_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
    "smsp__sass_average_branch_targets_threads_uniform.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "sm__cycles_active.avg", "sm__cycles_elapsed.avg.per_second",
]
STALL = "smsp__average_warps_issue_stalled_"


def num(d, k):
    return float(d[k][1].replace(",", "")) if k in d and d[k][1] not in ("", "n/a") else None


def scaled(d, k):
    """value in base units (ncu prints Kbyte / Mbyte / ms ... per row)"""
    v = num(d, k)
    if v is None:
        return None
    u = d[k][0].lower()
    for prefix, f in (("k", 1e3), ("m", 1e6), ("g", 1e9), ("t", 1e12)):
        if u.startswith(prefix + "byte"):
            return v * f
    return v * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}.get(u, 1.0)


def write_json(rep, out, units_per_launch):
    import json
    import os
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    d = dict(zip(rows[0], zip(rows[1], rows[2])))
    inst = num(d, "smsp__inst_executed.sum")
    lanes = num(d, "smsp__thread_inst_executed_per_inst_executed.ratio")
    rd, wr = scaled(d, "dram__bytes_read.sum"), scaled(d, "dram__bytes_write.sum")
    rec = {
        "source": os.path.basename(rep), "kernel": d.get("Kernel Name", ("", "?"))[1],
        "duration_s": scaled(d, "gpu__time_duration.sum"), "units_per_launch": units_per_launch,
        "registers_per_thread": num(d, "launch__registers_per_thread"),
        "issue_active_pct": num(d, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "pipe_alu_pct": num(d, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "pipe_fma_pct": num(d, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        "pipe_fmaheavy_cycles_pct": num(d, "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        "pipe_xu_pct": num(d, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        "pipe_lsu_pct": num(d, "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        "shared_pipe_pct": num(d, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
        "lanes_active_per_inst": lanes, "branch_targets_uniform_pct": num(d, "smsp__sass_average_branch_targets_threads_uniform.pct"),
        "warp_instructions": inst,
        "thread_instructions_per_unit": inst * lanes / units_per_launch if inst and lanes and units_per_launch else None,
        "dram_bytes_per_launch": (rd or 0.0) + (wr or 0.0) if rd is not None else None,
        "dram_throughput_pct": num(d, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    }
    with open(out, "w") as f:
        json.dump(rec, f, indent=1)
        f.write("\n")
    print(json.dumps(rec))


def main():
    rep = sys.argv[1]
    if len(sys.argv) > 3 and sys.argv[2] == "--json":
        return write_json(rep, sys.argv[3], float(sys.argv[4]) if len(sys.argv) > 4 else None)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, zip(units, vals)))
        print("=" * 100)
        print("kernel:", d.get("Kernel Name", ("", "?"))[1], " grid", d.get("Grid Size", ("", ""))[1], " block", d.get("Block Size", ("", ""))[1])
        for k in KEYS:
            if k in d:
                print(f"  {k:85s} {d[k][1]:>18s} {d[k][0]}")
        stalls = sorted(((float(v[1].replace(',', '')), k[len(STALL):-len('_per_issue_active.ratio')]) for k, v in d.items()
                         if k.startswith(STALL) and k.endswith("_per_issue_active.ratio")), reverse=True)
        print("  warp stall reasons (warps per issue-active cycle):", ", ".join(f"{n} {x:.2f}" for x, n in stalls[:8]))


if __name__ == "__main__":
    main()
