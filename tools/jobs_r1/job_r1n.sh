#!/bin/bash
mkdir -p gpurun_out
MAKB200_EIGH_TWOSTAGE=64 MAKB200_PROFILE=1 timeout 200 python tools/twostage_check.py 8192 2>&1 | grep -v "^\[makb200 profile\] stedc" | tail -24 | cut -c1-260 | tee gpurun_out/twostage_check.txt
