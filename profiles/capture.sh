#!/bin/sh
# ncu captures for profiles/ (run under gpurun on one B200; see B200_PROFILING.md).
# Usage: sh profiles/capture.sh <round-tag>
set -x
TAG=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --playouts-per-step 268435456"
# plain run first: must exit 0 before anything is profiled
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_$TAG.log; exit 1; }
# every launch with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
# the dominant kernel, full set
ncu --set full --clock-control none --import-source on -k regex:k_playout -s 3 -c 1 -o gpurun_out/prof_playout_$TAG $CMD > gpurun_out/ncu_playout_$TAG.log 2>&1
# the two other kernel families
# perft: launch 0 = the sharded depth-10 job, 1.. = depth 6 (x4), then depth 7 -- take a depth-7 launch
ncu --set full --clock-control none --import-source on -k regex:k_perft_tile -s 5 -c 1 -o gpurun_out/prof_perft_tile_$TAG $CMD > gpurun_out/ncu_perft_tile_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_canonical -s 1 -c 1 -o gpurun_out/prof_canon_$TAG $CMD > gpurun_out/ncu_canon_$TAG.log 2>&1
ls -la gpurun_out
