#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== gemm tests (TMA kernel default for aligned f64) =="
timeout 600 python -m pytest tests/test_gpu_gemm.py -q -x 2>&1 | tail -6
echo "== gemm probe =="
timeout 300 python tools/perf_probe.py gemm 2>&1 | grep "^gemm f64"
echo "== gemm probe, MAKB200_GEMM_SHORTK=0 (TMA kernel also for K <= 256) =="
MAKB200_GEMM_SHORTK=0 timeout 300 python tools/perf_probe.py gemm 2>&1 | grep "^gemm f64"
echo "== full suites touching GEMM =="
timeout 1200 python -m pytest tests/test_gpu_qr.py tests/test_gpu_eigh.py tests/test_gpu_svd_polar.py tests/test_gpu_tsqr.py tests/test_gpu_y_l1_shims.py -q 2>&1 | tail -5
echo "== bench C2 =="
MAKB200_GEMM_LOG=gpurun_out/gemm_log_c2c.txt timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | cut -c1-600
python tools/gemm_shapes.py gpurun_out/gemm_log_c2c.txt > gpurun_out/gemm_shapes_c2c.txt; head -18 gpurun_out/gemm_shapes_c2c.txt
echo "== bench C2, MAKB200_GEMM_SHORTK=0 =="
MAKB200_GEMM_SHORTK=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | cut -c1-330
} > gpurun_out/r2j.log 2>&1
tail -100 gpurun_out/r2j.log
