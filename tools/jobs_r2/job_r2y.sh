#!/bin/bash
set -u
mkdir -p gpurun_out
{ timeout 900 python tools/scale_probe2.py 2>&1 | tail -14; } > gpurun_out/r2y.log 2>&1
tail -40 gpurun_out/r2y.log
