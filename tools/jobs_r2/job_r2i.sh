#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== gemm tests =="
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_tsqr.py tests/test_gpu_eigh.py -q 2>&1 | tail -4
echo "== gemm probe (short-K config on) =="
timeout 300 python tools/perf_probe.py gemm 2>&1 | grep "^gemm"
echo "== gemm probe (MAKB200_GEMM_SHORTK=0) =="
MAKB200_GEMM_SHORTK=0 timeout 300 python tools/perf_probe.py gemm 2>&1 | grep "^gemm f64"
echo "== gemm probe (MAKB200_GEMM_VARIANT=3: 128x64 tiles everywhere) =="
MAKB200_GEMM_VARIANT=3 timeout 300 python tools/perf_probe.py gemm 2>&1 | grep "^gemm f64"
echo "== bench C2 =="
MAKB200_GEMM_LOG=gpurun_out/gemm_log_c2b.txt timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | cut -c1-600
python tools/gemm_shapes.py gpurun_out/gemm_log_c2b.txt | head -16
echo "== bench TSQR N=1 =="
MAKB200_PROFILE=1 timeout 600 python bench.py --workload tsqr --steps 3 --warmup 3 --no-cpu 2>&1 | tail -2 | cut -c1-500
echo "== ncu: gemm 8192^3 f64 =="
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_kernel -s 6 -c 1 -o gpurun_out/r2_gemm8192 -f python tools/perf_probe.py gemm > gpurun_out/ncu_gemm.log 2>&1
tail -2 gpurun_out/ncu_gemm.log
} > gpurun_out/r2i.log 2>&1
tail -100 gpurun_out/r2i.log
