#!/bin/bash
# final round-1 record: full GPU tests, both bench arms, full-size sweep with the final build, GEMM shape probe
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/clocks.csv &
SMI=$!
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref10.json 2> gpurun_out/bench10.err
python bench.py --steps 3 --warmup 3 > gpurun_out/bench10.json 2>> gpurun_out/bench10.err
kill $SMI
tail -1 gpurun_out/bench10.json | cut -c1-420
MAKB200_PROFILE=1 timeout 400 python tools/config_sweep.py C1 C2 C2c > gpurun_out/config_sweep3.jsonl 2> gpurun_out/config_sweep3.err
cut -c1-260 gpurun_out/config_sweep3.jsonl; grep -E "eigh:|svd:|polar:" gpurun_out/config_sweep3.err | cut -c1-220 | tail -6
timeout 200 python tools/perf_probe.py gemm 2>&1 | grep -E "^gemm" | tee gpurun_out/gemm_probe.txt
