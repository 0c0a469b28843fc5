#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== batched suites =="
timeout 1200 python -m pytest tests/test_gpu_eigh.py tests/test_gpu_svd_polar.py tests/test_gpu_y_rankdef.py tests/test_gpu_y_trunc.py tests/test_gpu_y_vals.py -q 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_x_config_size.py -q -k "ragged or config3 or blocks" 2>&1 | tail -3
echo "== full C3: qr + svdtrunc + eigh, all 20000 blocks, 1 GPU (lock-step QDWH + eigensolve + tail) =="
MAKB200_BENCH_BIG_CAP=100000 timeout 1500 python tools/batched_bench.py 20000 512 svdtrunc,eigh 2>&1 | grep -E "\"(svdtrunc|eigh)_|blocks\"|ms_max|blocks_per_s"
} > gpurun_out/r3c.log 2>&1
tail -70 gpurun_out/r3c.log
