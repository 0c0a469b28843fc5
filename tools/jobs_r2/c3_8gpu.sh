#!/bin/bash
# C3 (20 000 c128 blocks 16-512) LPT-partitioned over N GPUs, no data-path collective: all blocks of every bucket
set -u
mkdir -p gpurun_out
N=${1:-8}
{
nvidia-smi -L | head -8
MAKB200_BENCH_BIG_CAP=100000 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/batched_bench.py 20000 512 qr,svdtrunc,eigh 2>&1 | grep -E "workload|n_gpus|lpt_imbalance|\"(qr|svdtrunc|eigh)_|\"blocks\"|ms_max|blocks_per_s"
} > gpurun_out/c3_${N}gpu.log 2>&1
tail -80 gpurun_out/c3_${N}gpu.log
