#!/bin/bash
# Round-2 job C (1 GPU): persistent TMA tridiagonalisation column kernels (trd2.cuh): parity, then timing vs round-1 kernels.
set -u
mkdir -p gpurun_out
{
echo "== eigh tests (v2 kernels default) =="
timeout 600 python -m pytest tests/test_gpu_eigh.py tests/test_gpu_y_vals.py -x -q 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_x_config_size.py -x -q -k "eigh" 2>&1 | tail -8
echo "== timing eigh (v2) =="
MAKB200_PROFILE=1 timeout 300 python tools/perf_probe.py eigh 2>&1 | tail -24
echo "== timing eigh (round-1 kernels, MAKB200_SYMV_V2=0) =="
MAKB200_SYMV_V2=0 MAKB200_PROFILE=1 timeout 300 python tools/perf_probe.py eigh 2>&1 | grep -E "eigh_full|hetrd" | tail -12
} > gpurun_out/r2c.log 2>&1
tail -70 gpurun_out/r2c.log
