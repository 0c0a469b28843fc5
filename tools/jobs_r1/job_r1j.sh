#!/bin/bash
# PDL default on + L2 prefetch before the dependency wait: full GPU tests, A/B phase timing, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
for PF in 1 0; do
  echo "PREFETCH=$PF"; MAKB200_SYMV_PREFETCH=$PF MAKB200_PROFILE=1 timeout 300 python tools/config_sweep.py C2 2>&1 | grep -E "eigh:" | cut -c1-200 | tail -2
done
python bench.py --steps 3 --warmup 3 > gpurun_out/bench9.json 2> gpurun_out/bench9.err
tail -1 gpurun_out/bench9.json | cut -c1-420
