#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== bhetrd kernel time inside batched eigh (192 blocks per bucket), ncu durations =="
MAKB200_BENCH_BIG_CAP=192 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:bhetrd --csv --log-file gpurun_out/ncu_bhetrd.csv python tools/batched_bench.py 20000 512 eigh > gpurun_out/bb_eigh_ncu.log 2>&1
grep -v "^==" gpurun_out/ncu_bhetrd.csv | awk -F'","' 'NR>1{print $5, $(NF)}' | tail -12
echo "== same run without ncu: bucket times =="
MAKB200_BENCH_BIG_CAP=192 timeout 900 python tools/batched_bench.py 20000 512 eigh 2>&1 | grep -E "eigh_(65|129|257)|ms_max|blocks_per_s" | tail -9
} > gpurun_out/r3a.log 2>&1
tail -40 gpurun_out/r3a.log
