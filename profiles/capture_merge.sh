#!/bin/sh
# ncu captures of the merge kernels (BFS expansion at the largest frontier, canonical dedup insert).
# Usage: sh profiles/capture_merge.sh <round-tag>     (one B200; after a plain run exited 0)
TAG=${1:-r01g}
mkdir -p gpurun_out
python scripts/bfs_profile_case.py > gpurun_out/plain_merge_$TAG.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:k_bfs_expand_insert -s 10 -c 1 -o gpurun_out/prof_bfs_$TAG python scripts/bfs_profile_case.py > gpurun_out/ncu_bfs_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_dedup_insert -s 1 -c 1 -o gpurun_out/prof_dedup_$TAG python scripts/bfs_profile_case.py > gpurun_out/ncu_dedup_$TAG.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bfs_$TAG.csv python scripts/bfs_only.py > gpurun_out/ncu_bfs_launches_$TAG.log 2>&1
