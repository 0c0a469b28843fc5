#!/bin/bash
# Round-1 final validation: GPU tests, QDWH threshold effect at full size, pooled SVD/eigh, bench both arms.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
MAKB200_PROFILE=1 timeout 600 python tools/config_sweep.py C2 C5 > gpurun_out/config_sweep2.jsonl 2> gpurun_out/config_sweep2.err; cat gpurun_out/config_sweep2.jsonl | cut -c1-420
grep "qdwh steps\|svd:" gpurun_out/config_sweep2.err | tail -4
timeout 600 python tools/batched_bench.py 20000 512 svd,eigh > gpurun_out/bse_pool32.json 2> gpurun_out/bse_pool32.err
python - <<PY
import json
t=open("gpurun_out/bse_pool32.json").read(); d=json.loads(t[t.index("{"):])
print("pool32", {k:(v["blocks"], round(v["ms_max_over_ranks"],1), round(v["blocks_per_s"])) for k,v in d["buckets"].items()})
PY
MAKB200_BQR_MIN_DIM=65 timeout 600 python tools/batched_bench.py 20000 512 qr > gpurun_out/bq2_min65.json 2> gpurun_out/bq2_min65.err
python - <<PY
import json
t=open("gpurun_out/bq2_min65.json").read(); d=json.loads(t[t.index("{"):])
print("min65", {k:(v["blocks"], round(v["ms_max_over_ranks"],2), round(v["alg_GFLOPs"])) for k,v in d["buckets"].items()})
PY
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/clocks.csv &
SMI=$!
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref7.json 2> gpurun_out/bench7.err
python bench.py --steps 3 --warmup 3 > gpurun_out/bench7.json 2>> gpurun_out/bench7.err
kill $SMI
tail -1 gpurun_out/bench7.json | cut -c1-900
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
