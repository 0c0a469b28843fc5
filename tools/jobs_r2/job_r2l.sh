#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== gemm tests (complex TMA kernel) =="
timeout 600 python -m pytest tests/test_gpu_gemm.py -q -x 2>&1 | tail -3
timeout 300 python tools/perf_probe.py gemm 2>&1 | grep "^gemm c128"
echo "== suites =="
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
echo "== mid-size single-matrix probe =="
MAKB200_PROFILE=1 timeout 300 python tools/midsize_probe.py 2>&1 | grep -v "stedc:\|qdwh estimate" | tail -60
} > gpurun_out/r2l.log 2>&1
tail -120 gpurun_out/r2l.log
