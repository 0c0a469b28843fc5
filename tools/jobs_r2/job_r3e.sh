#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== lock-step tests with the per-chunk sigma_min estimate =="
timeout 600 python -m pytest tests/test_gpu_svd_polar.py tests/test_gpu_y_rankdef.py tests/test_gpu_y_trunc.py -q 2>&1 | tail -4
MAKB200_LOCKSTEP_VERBOSE=1 timeout 300 python tools/lockstep_smoke.py 2>&1 | grep -E "schedule|svd max" | head -8
echo "== full C3 svdtrunc (estimate on) =="
MAKB200_BENCH_BIG_CAP=100000 timeout 900 python tools/batched_bench.py 20000 512 svdtrunc 2>&1 | grep -E "\"svdtrunc_|ms_max|blocks_per_s"
echo "== 600 per bucket, MAKB200_LS_ESTIMATE=0 =="
MAKB200_LS_ESTIMATE=0 MAKB200_BENCH_BIG_CAP=600 timeout 600 python tools/batched_bench.py 20000 512 svd 2>&1 | grep -E "svd_(65|129|257)|blocks_per_s" | tail -6
echo "== 600 per bucket, estimate on =="
MAKB200_BENCH_BIG_CAP=600 timeout 600 python tools/batched_bench.py 20000 512 svd 2>&1 | grep -E "svd_(65|129|257)|blocks_per_s" | tail -6
} > gpurun_out/r3e.log 2>&1
tail -60 gpurun_out/r3e.log
