#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== mid-size probe, TMA GEMM on (default) =="
timeout 300 python tools/midsize_probe.py 2>&1 | grep -E "^svd|^eigh"
echo "== mid-size probe, MAKB200_GEMM_TMA=0 =="
MAKB200_GEMM_TMA=0 timeout 300 python tools/midsize_probe.py 2>&1 | grep -E "^svd|^eigh"
echo "== batched svd 65-512 (64 per bucket), default =="
timeout 600 python tools/batched_bench.py 4000 512 svd 2>&1 | grep -E "svd_(65|129|257)|blocks_per_s" | head -12
echo "== batched svd 65-512, MAKB200_GEMM_TMA=0 =="
MAKB200_GEMM_TMA=0 timeout 600 python tools/batched_bench.py 4000 512 svd 2>&1 | grep -E "svd_(65|129|257)|blocks_per_s" | head -12
} > gpurun_out/r2m.log 2>&1
tail -60 gpurun_out/r2m.log
