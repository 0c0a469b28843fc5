#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== lock-step tests, adaptive chunks =="
timeout 900 python -m pytest tests/test_gpu_eigh.py tests/test_gpu_svd_polar.py tests/test_gpu_y_rankdef.py tests/test_gpu_y_trunc.py -q 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_x_config_size.py -q -k "ragged or config3 or blocks" 2>&1 | tail -3
echo "== full C3 svdtrunc + eigh, adaptive chunks =="
MAKB200_BENCH_BIG_CAP=100000 timeout 1500 python tools/batched_bench.py 20000 512 svdtrunc,eigh 2>&1 | grep -E "\"(svdtrunc|eigh)_|ms_max|blocks_per_s"
echo "== svd big buckets (600 per bucket) with the 64x128 / 128x128 grouped GEMM tiles (MAKB200_GROUPED_SMALL=0) =="
MAKB200_GROUPED_SMALL=0 MAKB200_BENCH_BIG_CAP=600 timeout 900 python tools/batched_bench.py 20000 512 svd 2>&1 | grep -E "svd_(65|129|257)|blocks_per_s" | tail -6
echo "== same, default tiles =="
MAKB200_BENCH_BIG_CAP=600 timeout 900 python tools/batched_bench.py 20000 512 svd 2>&1 | grep -E "svd_(65|129|257)|blocks_per_s" | tail -6
} > gpurun_out/r3d.log 2>&1
tail -70 gpurun_out/r3d.log
