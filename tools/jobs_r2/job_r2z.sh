#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== scale probe (Jacobi SVD kernel: scale-safe convergence test) =="
timeout 600 python tools/scale_probe.py 111 2>&1 | tail -10 | cut -c1-260
echo "== lock-step tests + batched suites =="
timeout 900 python -m pytest tests/test_gpu_svd_polar.py tests/test_gpu_y_rankdef.py tests/test_gpu_y_trunc.py tests/test_gpu_eigh.py -q 2>&1 | tail -4
echo "== full C3: qr + svdtrunc + eigh, all 20000 blocks, 1 GPU (lock-step QDWH) =="
MAKB200_BENCH_BIG_CAP=100000 timeout 1500 python tools/batched_bench.py 20000 512 qr,svdtrunc,eigh 2>&1 | tail -120
} > gpurun_out/r2z.log 2>&1
tail -150 gpurun_out/r2z.log
