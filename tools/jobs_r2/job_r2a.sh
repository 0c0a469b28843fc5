#!/bin/bash
# Round-2 bring-up job (ONE gpurun call): everything written after the round-1 GPU budget was spent.
# Ordered by risk: (1) ungated kernels that already run in `pytest -m gpu`, (2) gated kernels without
# inter-CTA waits, (3) the fused Q2 slab kernel, (4) the persistent (spin-waiting) chase kernel LAST and
# under a short timeout, then the timings that decide which opt-in paths become defaults.
#   gpurun --timeout 2400 -- 'bash tools/jobs_r2/job_r2a.sh'      (log: gpurun_out/r2a.log)
set -u
mkdir -p gpurun_out
{
echo "== (0) cuSOLVER / cuBLAS bar (library calls, same box) =="
timeout 600 python tools/cusolver_bar.py gpurun_out/r2_cusolver_bar.json 2>&1 | tail -40
echo "== (1) ungated: values-only, leading-rank, batched truncation, projections, rank-deficient SVD =="
timeout 900 python -m pytest tests/test_gpu_y_vals.py tests/test_gpu_y_trunc.py tests/test_gpu_y_projections.py tests/test_gpu_y_rankdef.py -q 2>&1 | tail -15
echo "== timings: eigh_vals vs eigh_full, svd_vals vs svd_compact, svd_trunc r=1024 (8192 f64) =="
timeout 600 python tools/vals_time.py 8192 1024
echo "== (2) gated, no inter-CTA waits: panel-blocked warp QR, one-launch / single-CTA tridiagonalisation, lower-triangle stage 1 =="
MAKB200_BRINGUP=1 timeout 900 python -m pytest tests/test_gpu_zz_bringup.py -q -k "panel_blocked or one_launch or single_launch or lower_triangle or stage1_lookahead" 2>&1 | tail -8
echo "== (3) gated: fused Q2 slab kernel =="
MAKB200_BRINGUP=1 timeout 600 python -m pytest tests/test_gpu_zz_bringup.py -q -k "fused_q2" 2>&1 | tail -8
echo "== tiny-block QR (16-32 c128, x16 replicas): warp kernel vs panel-blocked warp kernel =="
timeout 300 python tools/batched_bench.py 20000 32 qr 2>&1 | grep -E "qr_16|blocks_per_s|hbm_frac" | head -6
MAKB200_BQR_WARP_BLK=1 timeout 300 python tools/batched_bench.py 20000 32 qr 2>&1 | grep -E "qr_16|blocks_per_s|hbm_frac" | head -6
echo "== batched svd_trunc!(truncrank(n_i/2)) of the 16-64 blocks: compact SVD + one truncation launch =="
timeout 300 python tools/batched_bench.py 20000 64 svd,svdtrunc 2>&1 | grep -E "svd(trunc)?_(16|33)|blocks_per_s" | head -12
echo "== batched eigh 65-512 c128 (64 blocks per bucket): pooled per-block hetrd vs one-launch tridiagonalisation =="
timeout 600 python tools/batched_bench.py 4000 512 eigh 2>&1 | grep -E "eigh_(65|129|257)|blocks_per_s" | head -12
MAKB200_BHETRD=1 timeout 600 python tools/batched_bench.py 4000 512 eigh 2>&1 | grep -E "eigh_(65|129|257)|blocks_per_s" | head -12
echo "== batched svd 65-512 c128 (64 blocks per bucket): 2n launches per block vs single-CTA tridiagonalisation inside eigh_t; then a wider pool =="
timeout 900 python tools/batched_bench.py 4000 512 svd 2>&1 | grep -E "svd_(65|129|257)|blocks_per_s" | head -12
MAKB200_BHETRD=2 timeout 900 python tools/batched_bench.py 4000 512 svd 2>&1 | grep -E "svd_(65|129|257)|blocks_per_s" | head -12
MAKB200_BHETRD=2 MAKB200_POOL_STREAMS=32 MAKB200_POOL_THREADS=8 timeout 900 python tools/batched_bench.py 4000 512 svd 2>&1 | grep -E "svd_(65|129|257)|blocks_per_s" | head -12
echo "== (4) gated: persistent bulge chasing (cooperative launch with progress counters) =="
MAKB200_BRINGUP=1 timeout 300 python -m pytest tests/test_gpu_zz_bringup.py -q -x -k "persistent_chase" 2>&1 | tail -8
echo "== chase: wavefront launches vs persistent (n=8192) =="
timeout 300 python tools/sbr_time.py 8192 64 32
MAKB200_CHASE_PERSISTENT=1 timeout 120 python tools/sbr_time.py 8192 64 32
for g in 8 16 32 64; do echo "-- persistent, grid cap $g"; MAKB200_CHASE_PERSISTENT=1 MAKB200_CHASE_GRID=$g timeout 120 python tools/sbr_time.py 8192 64; done
echo "== two-stage eigh 8192 f64: round-1 kernels / +persistent chase / +fused Q2 (g = 64, 32; cw = 64, 32) =="
for cfg in "" "MAKB200_SY2SB_LOWER=1" "MAKB200_SY2SB_LOOKAHEAD=1" "MAKB200_SY2SB_LOWER=1 MAKB200_SY2SB_LOOKAHEAD=1" "MAKB200_Q2_FUSED=1" "MAKB200_SY2SB_LOWER=1 MAKB200_Q2_FUSED=1" "MAKB200_CHASE_PERSISTENT=1" "MAKB200_CHASE_PERSISTENT=1 MAKB200_Q2_FUSED=1" \
           "MAKB200_CHASE_PERSISTENT=1 MAKB200_Q2_FUSED=1 MAKB200_Q2_G=32" \
           "MAKB200_CHASE_PERSISTENT=1 MAKB200_Q2_FUSED=1 MAKB200_Q2_G=32 MAKB200_Q2_CW=32" \
           "MAKB200_SY2SB_LOWER=1 MAKB200_SY2SB_LOOKAHEAD=1 MAKB200_CHASE_PERSISTENT=1 MAKB200_Q2_FUSED=1" \
           "MAKB200_EIGH_TWOSTAGE=32 MAKB200_SY2SB_LOWER=1 MAKB200_SY2SB_LOOKAHEAD=1 MAKB200_CHASE_PERSISTENT=1 MAKB200_Q2_FUSED=1"; do
  echo "-- $cfg"
  env SKIP_SMALL=1 MAKB200_EIGH_TWOSTAGE=64 MAKB200_PHASES=1 $cfg timeout 300 python tools/twostage_check.py 8192 2>&1 | tail -12
done
echo "== (5) the assembled two-stage path with both bring-up kernels through the test =="
MAKB200_BRINGUP=1 timeout 600 python -m pytest tests/test_gpu_zz_bringup.py -q -k "two_stage" 2>&1 | tail -6
} > gpurun_out/r2a.log 2>&1
tail -150 gpurun_out/r2a.log
