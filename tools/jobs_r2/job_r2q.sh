#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== svd / polar phases 8192 f64 =="
MAKB200_PROFILE=1 timeout 600 python tools/perf_probe.py svd 2>&1 | grep -E "n=8192|qdwh steps|polar:|svd:|eigh:" | tail -12
echo "== eigh phases 8192 f64 =="
MAKB200_PROFILE=1 timeout 300 python tools/lda_probe.py 8192 2>&1 | grep -E "lda=8192|hetrd" | tail -3
} > gpurun_out/r2q.log 2>&1
tail -40 gpurun_out/r2q.log
