#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== new tests: L1 shims, svd_full, tags; batched rank-deficient =="
timeout 900 python -m pytest tests/test_gpu_y_l1_shims.py tests/test_gpu_y_rankdef.py tests/test_gpu_orthnull.py -q 2>&1 | tail -12
echo "== eigh tests with the register w2 kernel =="
timeout 600 python -m pytest tests/test_gpu_eigh.py -q 2>&1 | tail -3
timeout 300 python tools/trd2_debug.py 2>&1 | tail -4
echo "== lda probe =="
MAKB200_PROFILE=1 timeout 600 python tools/lda_probe.py 8192 2>&1 | grep -E "lda=|hetrd" | awk 'NR%4==0 || /lda=/'
} > gpurun_out/r2f.log 2>&1
tail -60 gpurun_out/r2f.log
