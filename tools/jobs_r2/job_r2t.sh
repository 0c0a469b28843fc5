#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== eigh tests (column pass with warp-owned columns) =="
timeout 900 python -m pytest tests/test_gpu_eigh.py tests/test_gpu_y_vals.py -q -x 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_x_config_size.py -q -x -k "eigh" 2>&1 | tail -3
echo "== phases =="
MAKB200_PROFILE=1 timeout 600 python tools/perf_probe.py eigh big 2>&1 | grep -E "eigh_full|hetrd" | awk 'NR%5==0 || /eigh_full/'
echo "== bench =="
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | cut -c1-330
echo "== ncu symv2 =="
timeout 600 ncu --set full --import-source on --clock-control none -k regex:trd_symv2 -s 1000 -c 1 -o gpurun_out/r2_symv2b -f env SKIP_SMALL=1 python tools/twostage_check.py 8192 > gpurun_out/ncu_symv2b.log 2>&1
tail -2 gpurun_out/ncu_symv2b.log
} > gpurun_out/r2t.log 2>&1
tail -40 gpurun_out/r2t.log
