#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== lock-step batched QDWH: tests =="
timeout 600 python -m pytest tests/test_gpu_svd_polar.py -q -x -k "lockstep" 2>&1 | tail -15
echo "== batched suites =="
timeout 900 python -m pytest tests/test_gpu_svd_polar.py tests/test_gpu_y_rankdef.py tests/test_gpu_y_trunc.py -q -x 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_x_config_size.py -q -x -k "ragged or config3 or blocks" 2>&1 | tail -5
echo "== batched svd, 600 per big bucket: lock-step QDWH (default) =="
MAKB200_BENCH_BIG_CAP=600 timeout 900 python tools/batched_bench.py 20000 512 svd 2>&1 | grep -E "svd_(65|129|257)|blocks_per_s" | tail -6
echo "== same, MAKB200_SVD_LOCKSTEP=0 =="
MAKB200_SVD_LOCKSTEP=0 MAKB200_BENCH_BIG_CAP=600 timeout 900 python tools/batched_bench.py 20000 512 svd 2>&1 | grep -E "svd_(65|129|257)|blocks_per_s" | tail -6
} > gpurun_out/r2w.log 2>&1
tail -60 gpurun_out/r2w.log
