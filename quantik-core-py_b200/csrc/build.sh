#!/bin/sh
# Builds libquantik_b200.so in-tree for sm_100a (no GPU needed: nvcc cross-compiles).
set -e
cd "$(dirname "$0")"
OUT=../quantik_b200/_lib
mkdir -p "$OUT"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 --std=c++17 --threads 0 \
     -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v ${QK_NVCC_EXTRA} \
     --shared -o "$OUT/libquantik_b200.so" \
     qk_abi.cu qk_prims.cu qk_playout.cu qk_perft.cu qk_dedup.cu qk_peak.cu qk_guided.cu qk_solve.cu -lcudart_static 2>&1
