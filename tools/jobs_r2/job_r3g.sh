#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== batched suites on the final tree =="
timeout 1200 python -m pytest tests/test_gpu_eigh.py tests/test_gpu_svd_polar.py tests/test_gpu_y_rankdef.py tests/test_gpu_y_trunc.py tests/test_gpu_y_vals.py tests/test_gpu_qr.py -q 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_x_config_size.py -q -k "ragged or config3 or blocks" 2>&1 | tail -3
echo "== full C3: qr + svdtrunc + eigh, all 20000 blocks, 1 GPU, final tree =="
MAKB200_BENCH_BIG_CAP=100000 timeout 1500 python tools/batched_bench.py 20000 512 qr,svdtrunc,eigh 2>&1 | grep -E "\"(qr|svdtrunc|eigh)_|\"blocks\"|ms_max|blocks_per_s|alg_GFLOPs"
} > gpurun_out/r3g.log 2>&1
tail -100 gpurun_out/r3g.log
