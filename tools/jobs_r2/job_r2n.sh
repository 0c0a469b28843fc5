#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== batched tests (graph replay default) =="
timeout 900 python -m pytest tests/test_gpu_svd_polar.py tests/test_gpu_eigh.py tests/test_gpu_y_trunc.py tests/test_gpu_y_vals.py tests/test_gpu_y_rankdef.py -q -x 2>&1 | tail -5
timeout 1200 python -m pytest tests/test_gpu_x_config_size.py -q -x -k "batched" 2>&1 | tail -5
echo "== batched svd 65-512 (64 per bucket): graphs on =="
timeout 600 python tools/batched_bench.py 4000 512 svd 2>&1 | grep -E "svd_(65|129|257)|blocks_per_s" | tail -6
echo "== same, MAKB200_BATCH_GRAPHS=0 =="
MAKB200_BATCH_GRAPHS=0 timeout 600 python tools/batched_bench.py 4000 512 svd 2>&1 | grep -E "svd_(65|129|257)|blocks_per_s" | tail -6
echo "== batched svd, 600 blocks per big bucket: graphs on =="
MAKB200_BENCH_BIG_CAP=600 timeout 900 python tools/batched_bench.py 20000 512 svd 2>&1 | grep -E "svd_|blocks_per_s|\"ms" | tail -16
echo "== batched eigh, 600 per big bucket: graphs on =="
MAKB200_BENCH_BIG_CAP=600 timeout 900 python tools/batched_bench.py 20000 512 eigh 2>&1 | grep -E "eigh_|blocks_per_s" | tail -10
} > gpurun_out/r2n.log 2>&1
tail -80 gpurun_out/r2n.log
