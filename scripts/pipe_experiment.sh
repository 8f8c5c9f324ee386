#!/bin/sh
# machine-model measurements: per-class issue rates + ncu pipe utilisation of k_playout
python - <<'PY'
import sys, json
sys.path.insert(0, "quantik-core-py_b200")
import quantik_b200.batch as B
r = B.pipe_rates(0)
sms = 148
for k, v in r.items():
    print(f"{k:14s} {v/1e12:8.2f} Tlane-op/s  = {v/sms/1.965e9:6.1f} lanes/clk/SM")
json.dump(r, open("gpurun_out/pipe_rates.json", "w"))
PY
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio
ncu --metrics $M --clock-control none -k regex:k_playout -s 3 -c 1 --csv --log-file gpurun_out/playout_pipes_$1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --playouts-per-step 268435456 > gpurun_out/playout_pipes_$1.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(l for l in open("gpurun_out/playout_pipes_$1.csv") if l.startswith('"'))]
for r in rows[1:]:
    print(f"{r[-3][:75]:75s} {r[-1]:>16s} {r[-2]}")
PY
