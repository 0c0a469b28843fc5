#!/bin/bash
# 2-GPU job: multi-rank TSQR parity (C ABI + NCCL tree) and the N=2 bench line
set -u
mkdir -p gpurun_out
N=${1:-2}
{
nvidia-smi -L
echo "== torchrun pytest tests/test_gpu_tsqr_multi.py (N=$N) =="
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 -m pytest tests/test_gpu_tsqr_multi.py -q -x 2>&1 | tail -15
echo "== bench --gpus $N (TSQR strong scaling) =="
NCCL_DEBUG=WARN MAKB200_PROFILE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 2>&1 | grep -v "^\[makb200 profile\] cholqr\|W0\|^$" | tail -12 | cut -c1-3000
echo "== reference arm --gpus $N =="
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 1 --warmup 1 2>&1 | tail -2 | cut -c1-1500
} > gpurun_out/r2h_$N.log 2>&1
tail -60 gpurun_out/r2h_$N.log
