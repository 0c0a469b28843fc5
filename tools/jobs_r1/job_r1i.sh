#!/bin/bash
# programmatic dependent launch in hetrd (opt-in): parity tests and phase timing with and without
mkdir -p gpurun_out
MAKB200_PDL=1 timeout 400 python -m pytest tests/test_gpu_eigh.py tests/test_gpu_svd_polar.py -m gpu -x -q 2>&1 | tail -3
for PDL in 1 0; do
  echo "PDL=$PDL"; MAKB200_PDL=$PDL MAKB200_PROFILE=1 timeout 300 python tools/config_sweep.py C2 2>&1 | grep -E "eigh:|\"op\": \"eigh" | cut -c1-300 | tail -3
done
