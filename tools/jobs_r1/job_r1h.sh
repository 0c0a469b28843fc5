#!/bin/bash
# register-resident tiny-QR kernel (opt-in) vs the shared-memory one; lock-step threshold at 33
mkdir -p gpurun_out
MAKB200_BQR_WARP_REG=1 timeout 300 python -m pytest tests/test_gpu_qr.py -m gpu -x -q -k "batched" 2>&1 | tail -3
for REG in 0 1; do
  MAKB200_BQR_WARP_REG=$REG timeout 300 python tools/batched_bench.py 20000 64 qr > gpurun_out/bq_reg$REG.json 2> gpurun_out/bq_reg$REG.err
  python - <<PY
import json
t=open("gpurun_out/bq_reg$REG.json").read(); d=json.loads(t[t.index("{"):])
print("REG=$REG", {k:(v["blocks"], round(v["ms_max_over_ranks"],3), round(v["hbm_frac_per_gpu"],4)) for k,v in d["buckets"].items()})
PY
done
MAKB200_BQR_MIN_DIM=33 timeout 300 python tools/batched_bench.py 20000 64 qr > gpurun_out/bq_min33.json 2> gpurun_out/bq_min33.err
python - <<PY
import json
t=open("gpurun_out/bq_min33.json").read(); d=json.loads(t[t.index("{"):])
print("MIN_DIM=33", {k:(v["blocks"], round(v["ms_max_over_ranks"],3)) for k,v in d["buckets"].items()})
PY
