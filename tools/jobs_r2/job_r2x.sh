#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== scale probe =="
timeout 600 python tools/scale_probe.py 111 2>&1 | tail -12
echo "== lock-step tests (all variants) =="
timeout 600 python -m pytest tests/test_gpu_svd_polar.py -q -k "lockstep" 2>&1 | grep -E "passed|failed|FAILED|Error" | head
} > gpurun_out/r2x.log 2>&1
tail -40 gpurun_out/r2x.log
