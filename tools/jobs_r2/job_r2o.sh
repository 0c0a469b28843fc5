#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== batched eigh, 600 per big bucket: default pool =="
MAKB200_BENCH_BIG_CAP=600 timeout 900 python tools/batched_bench.py 20000 512 eigh 2>&1 | grep -E "eigh_(65|129|257)|blocks_per_s" | tail -6
echo "== batched eigh, 600 per big bucket: MAKB200_BHETRD=1 (one-launch tridiagonalisation of all blocks) =="
MAKB200_BHETRD=1 MAKB200_BENCH_BIG_CAP=600 timeout 900 python tools/batched_bench.py 20000 512 eigh 2>&1 | grep -E "eigh_(65|129|257)|blocks_per_s" | tail -6
echo "== same, 2000 per bucket =="
MAKB200_BHETRD=1 MAKB200_BENCH_BIG_CAP=2000 timeout 900 python tools/batched_bench.py 20000 512 eigh 2>&1 | grep -E "eigh_(65|129|257)|blocks_per_s" | tail -6
echo "== graph replay test =="
timeout 600 python -m pytest tests/test_gpu_svd_polar.py -q -k graph 2>&1 | tail -3
} > gpurun_out/r2o.log 2>&1
tail -40 gpurun_out/r2o.log
