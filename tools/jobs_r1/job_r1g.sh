#!/bin/bash
# trd_w parallelism change + probe split-K: tests, per-kernel times of hetrd for three row-segment sizes, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
for SEG in 1024 2048 512; do
  echo "SEG=$SEG"; MAKB200_TRD_SEG=$SEG KTIME=1 REPS=1 timeout 300 python tools/prof_run.py eigh 8192 2>&1 | grep ktime
done
MAKB200_PROFILE=1 timeout 300 python tools/config_sweep.py C2 2>&1 | grep -E "eigh:|polar:|svd:|\"case\"" | cut -c1-330 | tail -8
python bench.py --steps 3 --warmup 3 > gpurun_out/bench8.json 2> gpurun_out/bench8.err
tail -1 gpurun_out/bench8.json | cut -c1-700
