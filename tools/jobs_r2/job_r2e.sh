#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== small-n sweep, persistent kernels =="
timeout 300 python tools/trd2_debug.py 2>&1 | tail -8
echo "== eigh / svd / rankdef / orthnull tests =="
timeout 900 python -m pytest tests/test_gpu_eigh.py tests/test_gpu_y_vals.py tests/test_gpu_y_rankdef.py tests/test_gpu_svd_polar.py tests/test_gpu_orthnull.py tests/test_gpu_qr.py -q 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_x_config_size.py -x -q -k "eigh or svd" 2>&1 | tail -4
echo "== timing eigh (v2) =="
MAKB200_PROFILE=1 timeout 300 python tools/perf_probe.py eigh big 2>&1 | grep -E "eigh_full|hetrd|resid" 
echo "== bench =="
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | cut -c1-2500
} > gpurun_out/r2e.log 2>&1
tail -80 gpurun_out/r2e.log
