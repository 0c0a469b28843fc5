#!/bin/bash
mkdir -p gpurun_out
for cfg in "64 32" "32 32" "48 48" "32 16" "64 16"; do
  set -- $cfg
  echo "b=$1 g=$2"
  SKIP_SMALL=1 MAKB200_EIGH_TWOSTAGE=$1 MAKB200_Q2_G=$2 MAKB200_PROFILE=1 timeout 100 python tools/twostage_check.py 8192 2>&1 | grep -E "eigh:|eigh_full" | tail -2 | cut -c1-220
done | tee gpurun_out/twostage_sweep.txt
