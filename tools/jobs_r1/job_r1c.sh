#!/bin/bash
# Round-1 session-3 validation job (run under gpurun): GPU tests, two-level lock-step batched QR,
# threaded stream pool for batched SVD/eigh, full-size config sweep.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
MAKB200_BQR_NBO=128 timeout 300 python -m pytest tests/test_gpu_qr.py -m gpu -x -q -k "batched" 2>&1 | tail -3
MAKB200_BQR_NBO=32 timeout 300 python -m pytest tests/test_gpu_qr.py -m gpu -x -q -k "batched" 2>&1 | tail -3
for NBO in 64 128 32; do
  MAKB200_BQR_NBO=$NBO timeout 600 python tools/batched_bench.py 20000 512 qr > gpurun_out/bq2_nbo$NBO.json 2> gpurun_out/bq2_nbo$NBO.err
  python - <<PY
import json
t=open("gpurun_out/bq2_nbo$NBO.json").read(); d=json.loads(t[t.index("{"):])
print("NBO=$NBO", {k:(v["blocks"], round(v["ms_max_over_ranks"],2), round(v["alg_GFLOPs"])) for k,v in d["buckets"].items()})
PY
done
for TH in 8 1; do
  MAKB200_POOL_THREADS=$TH timeout 900 python tools/batched_bench.py 20000 512 svd,eigh > gpurun_out/bse_th$TH.json 2> gpurun_out/bse_th$TH.err
  python - <<PY
import json
t=open("gpurun_out/bse_th$TH.json").read(); d=json.loads(t[t.index("{"):])
print("threads=$TH", {k:(v["blocks"], round(v["ms_max_over_ranks"],1), round(v["blocks_per_s"])) for k,v in d["buckets"].items()})
PY
done
timeout 900 python tools/config_sweep.py > gpurun_out/config_sweep.jsonl 2> gpurun_out/config_sweep.err; cat gpurun_out/config_sweep.jsonl; tail -3 gpurun_out/config_sweep.err
