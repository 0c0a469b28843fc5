#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== smoke =="
timeout 300 python tools/lockstep_smoke.py 2>&1 | tail -12
echo "== compute-sanitizer memcheck on the smoke =="
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python tools/lockstep_smoke.py 2>&1 | grep -E "ERROR SUMMARY|Invalid|at .*\(|max dw|max dS" | head -30
echo "== lock-step tests =="
timeout 900 python -m pytest tests/test_gpu_eigh.py tests/test_gpu_svd_polar.py -q -k "lockstep" 2>&1 | tail -15
} > gpurun_out/r3b.log 2>&1
tail -60 gpurun_out/r3b.log
