#!/bin/bash
# Round profile job (run under gpurun): bench line, FP64 peak, ncu launch list of the bench command,
# ncu --set full captures of the two dominant kernels. Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
python tools/peak_fp64.py > gpurun_out/peak.log 2>&1
cp gpurun_out/fp64_peak.json profiles/fp64_peak.json 2>/dev/null
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > gpurun_out/clocks.csv &
SMI=$!
python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
kill $SMI
# launch list of the same command (reduced warm-up under ncu; shares, not absolutes)
MAKB200_BENCH_UNDER_NCU=1 timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 120000 --csv \
   --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/launches.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 2:]:
    if len(r) < 5: continue
    name = r[4].split("(")[0]
    try: v = float(r[-1].replace(",", ""))
    except ValueError: continue
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
with open("gpurun_out/launches_summary.txt", "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 1 --warmup 0 --no-cpu\n")
    f.write("(per-launch times are cold-cache and serialised: compare SHARES)\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        f.write(f"{k[:80]:80s} n={v[0]:7d} total={v[1]/1e6:10.3f} ms  {100*v[1]/tot:5.1f}%\n")
    f.write(f"total {tot/1e6:.3f} ms over {sum(v[0] for v in agg.values())} launches\n")
print(open("gpurun_out/launches_summary.txt").read())
PY
gzip -f gpurun_out/launches.csv
# full captures (3 launches each, mid-run)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trd_dots -s 1500 -c 3 -o gpurun_out/prof_dots \
   python tools/prof_run.py eigh 8192 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 40 -c 3 -o gpurun_out/prof_gemm \
   python tools/prof_run.py qr 8192 > /dev/null 2>&1
for f in dots gemm; do
  ncu -i gpurun_out/prof_$f.ncu-rep --page raw --csv 2>/dev/null | python - "$f" <<'PY'
import csv, sys
rows = list(csv.reader(sys.stdin))
if len(rows) > 2:
    hdr = rows[0]
    want = ["Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "smsp__inst_executed_pipe_fp64_op_dmma.sum",
            "sm__inst_executed_pipe_tensor_op_dmma.sum", "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
    idx = [(h, i) for i, h in enumerate(hdr) if h in want]
    with open(f"gpurun_out/prof_{sys.argv[1]}_summary.txt", "w") as f:
        for r in rows[2:]:
            f.write(" | ".join(f"{h}={r[i]}" for h, i in idx) + "\n")
        f.write("units: " + " | ".join(f"{h}={rows[1][i]}" for h, i in idx) + "\n")
    print(open(f"gpurun_out/prof_{sys.argv[1]}_summary.txt").read())
PY
done
cat gpurun_out/bench.json gpurun_out/bench_ref.json
tail -3 gpurun_out/bench.err
