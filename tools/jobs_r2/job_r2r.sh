#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== polar/svd tests with Cholesky look-ahead =="
timeout 900 python -m pytest tests/test_gpu_svd_polar.py tests/test_gpu_y_rankdef.py tests/test_gpu_tsqr.py -q -x 2>&1 | tail -3
timeout 1200 python -m pytest tests/test_gpu_x_config_size.py -q -x -k "polar or svd_compact" 2>&1 | tail -3
echo "== phases =="
MAKB200_PROFILE=1 timeout 600 python tools/perf_probe.py svd 2>&1 | grep -E "n=8192|qdwh steps|polar:|svd:" | tail -5
echo "== look-ahead off =="
MAKB200_POTRF_LOOKAHEAD=0 MAKB200_PROFILE=1 timeout 600 python tools/perf_probe.py svd 2>&1 | grep -E "n=8192|qdwh steps" | tail -3
} > gpurun_out/r2r.log 2>&1
tail -40 gpurun_out/r2r.log
