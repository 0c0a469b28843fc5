#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== C1 qr 4096: phases + GEMM shapes =="
MAKB200_PROFILE=1 MAKB200_GEMM_LOG=gpurun_out/gemm_log_c1b.txt timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --ops qr --n 4096 2>&1 | tail -4 | cut -c1-400
python tools/gemm_shapes.py gpurun_out/gemm_log_c1b.txt | head -30
echo "== launch list (ncu gpu__time_duration) of one qr_compact 4096 =="
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/qr_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --ops qr --n 4096 > /dev/null 2>&1
python tools/ncu_summ.py gpurun_out/qr_launches.csv | head -30
} > gpurun_out/r2u.log 2>&1
tail -80 gpurun_out/r2u.log
