#!/bin/bash
# 8-GPU evidence for the batched config (BASELINE configs[2]): LPT block partition, no data-path collective.
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 \
   tools/batched_bench.py 20000 512 qr,svd > gpurun_out/batched_N8b.json 2> gpurun_out/batched_N8b.err
python - <<'PY'
import json
t=open("gpurun_out/batched_N8b.json").read(); d=json.loads(t[t.index("{"):])
print("N=8 imbalance", d["lpt_imbalance"], {k:(v["blocks"], round(v["ms_max_over_ranks"],2), round(v["blocks_per_s"])) for k,v in d["buckets"].items()})
PY
tail -3 gpurun_out/batched_N8b.err
