#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== gemm tests (staged C tile for K <= 1024) =="
timeout 600 python -m pytest tests/test_gpu_gemm.py -q -x 2>&1 | tail -3
echo "== suites touching GEMM =="
timeout 1500 python -m pytest tests/test_gpu_qr.py tests/test_gpu_eigh.py tests/test_gpu_svd_polar.py tests/test_gpu_tsqr.py tests/test_gpu_y_l1_shims.py tests/test_gpu_y_rankdef.py -q -x 2>&1 | tail -3
echo "== bench C2 + shapes =="
MAKB200_GEMM_LOG=gpurun_out/gemm_log_c2f.txt timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | cut -c1-330
python tools/gemm_shapes.py gpurun_out/gemm_log_c2f.txt > gpurun_out/gemm_shapes_c2f.txt; head -22 gpurun_out/gemm_shapes_c2f.txt
echo "== bench C2, MAKB200_GEMM_CSTAGE_MAXK=0 =="
MAKB200_GEMM_CSTAGE_MAXK=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | cut -c1-330
echo "== C1 =="
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --ops qr --n 4096 2>&1 | tail -1 | cut -c1-330
} > gpurun_out/r2v.log 2>&1
tail -60 gpurun_out/r2v.log
