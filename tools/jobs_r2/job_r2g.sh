#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== L1 shims after the orgqr fix =="
timeout 900 python -m pytest tests/test_gpu_y_l1_shims.py -q 2>&1 | tail -5
echo "== GEMM shape log of one bench step (eigh+svd 8192 f64) =="
MAKB200_GEMM_LOG=gpurun_out/gemm_log_c2.txt timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu 2>&1 | tail -1 | cut -c1-400
python tools/gemm_shapes.py gpurun_out/gemm_log_c2.txt
echo "== same for qr_compact 4096 (C1) =="
MAKB200_GEMM_LOG=gpurun_out/gemm_log_c1.txt timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --ops qr --n 4096 2>&1 | tail -1 | cut -c1-300
python tools/gemm_shapes.py gpurun_out/gemm_log_c1.txt
} > gpurun_out/r2g.log 2>&1
tail -120 gpurun_out/r2g.log
