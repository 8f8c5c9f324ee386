#!/bin/sh
# Round-2 ncu captures (one B200, under gpurun; see B200_PROFILING.md).  Usage: sh profiles/capture_r02.sh <tag>
# Each capture only after the plain command has exited 0; numbers printed under ncu are never bench values.
TAG=${1:-r02h}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --playouts-per-step 268435456"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_$TAG.log; exit 1; }
# every launch with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
# the dominant kernel, full set
ncu --set full --clock-control none --import-source on -k regex:k_playout_single -s 3 -c 1 -o gpurun_out/prof_playout_$TAG $CMD > gpurun_out/ncu_playout_$TAG.log 2>&1
# perft: launches 0-3 = the sharded depth-10 job (warm-up + 3 runs), 4-7 = depth 6 (x4), 8.. = depth 7
ncu --set full --clock-control none --import-source on -k regex:k_perft_tile -s 8 -c 1 -o gpurun_out/prof_perft_tile_$TAG $CMD > gpurun_out/ncu_perft_tile_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_canonical -s 1 -c 1 -o gpurun_out/prof_canon_$TAG $CMD > gpurun_out/ncu_canon_$TAG.log 2>&1
# the merge kernels: BFS expansion at the largest frontier, two-level dedup (routing pass + one part's insert)
python scripts/bfs_profile_case.py > gpurun_out/plain_merge_$TAG.log 2>&1 || { echo "plain merge run failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:k_bfs_expand_insert -s 10 -c 1 -o gpurun_out/prof_bfs_$TAG python scripts/bfs_profile_case.py > gpurun_out/ncu_bfs_$TAG.log 2>&1
ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_bucket_states -s 1 -c 1 -o gpurun_out/prof_bucket_$TAG python scripts/bfs_profile_case.py > gpurun_out/ncu_bucket_$TAG.log 2>&1
ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_part_insert -s 12 -c 1 -o gpurun_out/prof_part_insert_$TAG python scripts/bfs_profile_case.py > gpurun_out/ncu_part_insert_$TAG.log 2>&1
# guided rollouts
ncu --set full --clock-control none --import-source on -k regex:k_playout_guided -s 1 -c 1 -o gpurun_out/prof_guided_$TAG python scripts/guided_case.py > gpurun_out/ncu_guided_$TAG.log 2>&1
ls -la gpurun_out | grep $TAG
