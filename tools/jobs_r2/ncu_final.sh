#!/bin/bash
# ncu --set full of the two dominant kernels on the final tree (3 launches each, warm)
set -u
mkdir -p gpurun_out
{
# gemm_tma_kernel<double>: 8192^3 (launches 8..10 of the probe = the timed 8192^3 repetitions)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tma_kernel -s 9 -c 2 -f -o gpurun_out/prof_gemm_tma python tools/perf_probe.py gemm 2>&1 | grep -E "gemm f64 (4096x4096x4096|8192x8192x8192)|PROF" | head -5
# trd_symv2_kernel<double>: columns ~1000 of an 8192^2 eigh
REPS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:trd_symv2 -s 1000 -c 2 -f -o gpurun_out/prof_symv2 python tools/prof_run.py eigh 8192 2>&1 | tail -2
ls -la gpurun_out/*.ncu-rep
} > gpurun_out/ncu_final.log 2>&1
tail -12 gpurun_out/ncu_final.log
