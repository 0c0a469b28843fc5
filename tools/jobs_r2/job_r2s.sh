#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== polar/svd/tsqr tests (new triangular inverse in the Cholesky block kernel) =="
timeout 900 python -m pytest tests/test_gpu_svd_polar.py tests/test_gpu_y_rankdef.py tests/test_gpu_tsqr.py tests/test_gpu_y_trunc.py -q -x 2>&1 | tail -3
timeout 1200 python -m pytest tests/test_gpu_x_config_size.py -q -x -k "polar or svd_compact" 2>&1 | tail -3
echo "== phases =="
MAKB200_PROFILE=1 timeout 600 python tools/perf_probe.py svd 2>&1 | grep -E "n=8192|qdwh steps|polar:|svd:" | tail -5
echo "== bench =="
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | cut -c1-330
echo "== tsqr N=1 =="
MAKB200_PROFILE=1 timeout 600 python bench.py --workload tsqr --steps 3 --warmup 3 --no-cpu 2>&1 | tail -2 | cut -c1-330
} > gpurun_out/r2s.log 2>&1
tail -40 gpurun_out/r2s.log
