#!/bin/bash
# Multi-GPU evidence for the configs that shard (BASELINE configs[2] batched, configs[3] TSQR).
# Run under: gpurun --gpus 8 -- bash tools/multi_gpu_job.sh
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for N in 2 4 8; do
  $TR --nproc-per-node $N --master-port $((29600+N)) tools/tsqr_multi.py 2>&1 | grep -E "^\{|parity" > gpurun_out/tsqr_N$N.log
  tail -2 gpurun_out/tsqr_N$N.log
done
python tools/tsqr_multi.py 2>&1 | grep -E "^\{" > gpurun_out/tsqr_N1.log; tail -1 gpurun_out/tsqr_N1.log
python tools/batched_bench.py 20000 512 qr,svd > gpurun_out/batched_N1.json 2>gpurun_out/batched_N1.err
for N in 2 8; do
  $TR --nproc-per-node $N --master-port $((29700+N)) tools/batched_bench.py 20000 512 qr,svd > gpurun_out/batched_N$N.json 2> gpurun_out/batched_N$N.err
done
python - <<'PY'
import json
for N in (1, 2, 8):
    try:
        txt = open(f"gpurun_out/batched_N{N}.json").read()
        d = json.loads(txt[txt.index("{"):])
        tot = {}
        for k, v in d["buckets"].items():
            op = k.split("_")[0]
            tot.setdefault(op, [0, 0.0])
            tot[op][0] += v["blocks"]; tot[op][1] += v["ms_max_over_ranks"]
        print("batched N=", N, "imbalance", round(d["lpt_imbalance"], 3), {op: (b, round(ms, 1)) for op, (b, ms) in tot.items()})
    except Exception as e:
        print("batched N=", N, "ERR", e)
PY
