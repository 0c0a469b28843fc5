#!/bin/bash
set -u
mkdir -p gpurun_out
{
for C in 192 296 444; do
echo "== svd 257-512, all 3991 blocks, MAKB200_SVD_CHUNK_MIN=$C =="
MAKB200_SVD_CHUNK_MIN=$C MAKB200_BENCH_BIG_CAP=100000 timeout 600 python tools/batched_bench.py 20000 512 svd 2>&1 | grep -E "svd_(129|257)|ms_max" | tail -4
done
} > gpurun_out/r3f.log 2>&1
tail -30 gpurun_out/r3f.log
