#!/bin/bash
mkdir -p gpurun_out
for SM in 0 1; do
  echo "GROUPED_SMALL=$SM b=64 g=64"
  SKIP_SMALL=1 MAKB200_GROUPED_SMALL=$SM MAKB200_EIGH_TWOSTAGE=64 MAKB200_PROFILE=1 timeout 100 python tools/twostage_check.py 8192 2>&1 | grep -E "eigh:|eigh_full" | tail -2 | cut -c1-220
done | tee gpurun_out/twostage_grouped.txt
