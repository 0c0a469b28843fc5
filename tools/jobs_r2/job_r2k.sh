#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== gemm tests =="
timeout 600 python -m pytest tests/test_gpu_gemm.py -q -x 2>&1 | tail -3
echo "== bench C2 default (short-K: two-CTA cp.async config with batched epilogue loads) =="
MAKB200_GEMM_LOG=gpurun_out/gemm_log_c2d.txt timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | cut -c1-330
python tools/gemm_shapes.py gpurun_out/gemm_log_c2d.txt | grep -E "launches| 128 +[0-9]+ +[0-9.]+ +[0-9.]+%"
echo "== bench C2 MAKB200_GEMM_SHORTK=0 (TMA kernel for short K too) =="
MAKB200_GEMM_SHORTK=0 MAKB200_GEMM_LOG=gpurun_out/gemm_log_c2e.txt timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | cut -c1-330
python tools/gemm_shapes.py gpurun_out/gemm_log_c2e.txt | grep -E "launches| 128 +[0-9]+ +[0-9.]+ +[0-9.]+%"
echo "== bench TSQR N=1 =="
MAKB200_PROFILE=1 timeout 600 python bench.py --workload tsqr --steps 3 --warmup 3 --no-cpu 2>&1 | tail -2 | cut -c1-400
echo "== C1 qr 4096 =="
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --ops qr --n 4096 2>&1 | tail -1 | cut -c1-330
echo "== C3 CPU baseline (sample 120 per bucket) =="
timeout 900 python tools/c3_cpu_baseline.py 120 gpurun_out/r2_c3_cpu_baseline.json 2>&1 | tail -12
} > gpurun_out/r2k.log 2>&1
tail -100 gpurun_out/r2k.log
