#!/bin/bash
# bring-up of the experimental band -> tridiagonal chase kernel: parity at small sizes, timing at 8192
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sbr.py -m gpu -x -q 2>&1 | tail -6
timeout 200 python tools/sbr_time.py 2048 64 2>&1 | tail -3
timeout 200 python tools/sbr_time.py 8192 64 32 2>&1 | tail -5 | tee gpurun_out/sbr_time.txt
