#!/bin/bash
# Round-2 job B (1 GPU): full GPU test suite incl. the config-size parity tests, bench N=1 (C2), TSQR at 1 GPU.
set -u
mkdir -p gpurun_out
{
echo "== pytest -m gpu (all) =="
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== bench N=1 (C2) =="
timeout 600 python bench.py --steps 3 --warmup 3 2>&1 | tail -3
echo "== bench --workload tsqr N=1 =="
MAKB200_PROFILE=1 timeout 600 python bench.py --workload tsqr --steps 3 --warmup 3 2>&1 | tail -8
} > gpurun_out/r2b.log 2>&1
tail -60 gpurun_out/r2b.log
