#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== small-n sweep, persistent kernels =="
timeout 300 python tools/trd2_debug.py 2>&1 | tail -30
echo "== small-n sweep, round-1 kernels =="
MAKB200_SYMV_V2=0 timeout 300 python tools/trd2_debug.py 2>&1 | tail -5
echo "== eigh tests =="
timeout 600 python -m pytest tests/test_gpu_eigh.py tests/test_gpu_y_vals.py tests/test_gpu_y_rankdef.py -q 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_x_config_size.py -x -q -k "eigh" 2>&1 | tail -4
echo "== timing eigh (v2) =="
MAKB200_PROFILE=1 timeout 300 python tools/perf_probe.py eigh 2>&1 | grep -E "eigh_full|hetrd|resid" | tail -12
echo "== ncu: one symv2 launch + one w2 launch deep in a 8192 f64 run =="
timeout 600 ncu --set full --import-source on --clock-control none -k regex:trd_symv2 -s 1000 -c 2 -o gpurun_out/r2_symv2 -f env SKIP_SMALL=1 python tools/twostage_check.py 8192 > gpurun_out/ncu_symv2.log 2>&1
tail -3 gpurun_out/ncu_symv2.log
timeout 600 ncu --set full --clock-control none -k regex:trd_w2 -s 1000 -c 2 -o gpurun_out/r2_w2 -f env SKIP_SMALL=1 python tools/twostage_check.py 8192 > gpurun_out/ncu_w2.log 2>&1
tail -3 gpurun_out/ncu_w2.log
} > gpurun_out/r2d.log 2>&1
tail -80 gpurun_out/r2d.log
