#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== batched tests (phased SVD + one-launch tridiagonalisation default) =="
timeout 900 python -m pytest tests/test_gpu_svd_polar.py tests/test_gpu_eigh.py tests/test_gpu_y_trunc.py tests/test_gpu_y_vals.py tests/test_gpu_y_rankdef.py tests/test_gpu_zz_bringup.py -q -x 2>&1 | tail -5
timeout 1200 python -m pytest tests/test_gpu_x_config_size.py -q -x -k "batched" 2>&1 | tail -3
echo "== batched svd, 600 per big bucket: phased =="
MAKB200_BENCH_BIG_CAP=600 timeout 900 python tools/batched_bench.py 20000 512 svd 2>&1 | grep -E "svd_(65|129|257)|blocks_per_s" | tail -6
echo "== batched svd, 600 per big bucket: MAKB200_SVD_PHASED=0 =="
MAKB200_SVD_PHASED=0 MAKB200_BENCH_BIG_CAP=600 timeout 900 python tools/batched_bench.py 20000 512 svd 2>&1 | grep -E "svd_(65|129|257)|blocks_per_s" | tail -6
echo "== full C3: qr + svdtrunc + eigh, all 20000 blocks, 1 GPU =="
MAKB200_BENCH_BIG_CAP=100000 timeout 1500 python tools/batched_bench.py 20000 512 qr,svdtrunc,eigh 2>&1 | tail -80
} > gpurun_out/r2p.log 2>&1
tail -130 gpurun_out/r2p.log
