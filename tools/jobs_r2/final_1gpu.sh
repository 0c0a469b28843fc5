#!/bin/bash
# Round-2 final single-GPU evidence: full GPU suite, smoke, bench line + reference arm, ncu launch list and DRAM traffic.
set -u
mkdir -p gpurun_out
{
echo "== pytest -m gpu =="
timeout 2400 python -m pytest tests/ -q -m gpu 2>&1 | tail -6
echo "== smoke =="
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== bench (default flags) =="
timeout 1500 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -1 gpurun_out/bench_final.json | cut -c1-3000
echo "== bench --impl reference =="
timeout 1500 python bench.py --impl reference > gpurun_out/bench_ref_final.json 2>> gpurun_out/bench_final.err; tail -1 gpurun_out/bench_ref_final.json | cut -c1-1500
echo "== C1 (qr_compact 4096) and c128 lines =="
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --ops qr --n 4096 2>&1 | tail -1 | cut -c1-400
echo "== ncu launch list (first 6000 launches of one step) =="
MAKB200_BENCH_UNDER_NCU=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/launches_r2.csv") if l.startswith('"')))
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if len(r) < 5: continue
    name = r[4].split("(")[0]
    try: v = float(r[-1].replace(",", ""))
    except ValueError: continue
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
with open("gpurun_out/launches_r2_summary.txt", "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 6000  python bench.py --steps 1 --warmup 0 --no-cpu\n")
    f.write("(first 6000 launches of the step = the start of eigh_full!'s tridiagonalisation; per-launch times are cold-cache and serialised: compare SHARES)\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
        f.write(f"{k[:90]:90s} n={v[0]:7d} total={v[1]/1e6:10.3f} ms  {100*v[1]/max(tot,1):5.1f}%\n")
    f.write(f"total {tot/1e6:.3f} ms over {sum(v[0] for v in agg.values())} launches\n")
print(open("gpurun_out/launches_r2_summary.txt").read())
PY
gzip -f gpurun_out/launches_r2.csv
echo "== ncu DRAM traffic of the column kernel and the TMA GEMM (first 400 matching launches) =="
MAKB200_BENCH_UNDER_NCU=1 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"trd_symv2|gemm_tma" -c 400 --csv --log-file gpurun_out/traffic_r2.csv python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/bench_under_ncu2.log 2>&1
python tools/traffic_summ.py gpurun_out/traffic_r2.csv gpurun_out/traffic_r2.json 8192 "first 400 launches of trd_symv2 / gemm_tma in the step (columns 0.. of the first tridiagonalisation and its her2k updates)" | tail -30
} > gpurun_out/final_1gpu.log 2>&1
tail -120 gpurun_out/final_1gpu.log
