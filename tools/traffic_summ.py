"""Turn an ncu --csv capture with dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum
(per launch, of `python bench.py --steps 1 --warmup 0 --no-cpu`) into profiles/traffic.json:
per kernel class the launch count, mean DRAM bytes per launch and total device time.
bench.py reads that file to fill roofline.traffic.
  python tools/traffic_summ.py gpurun_out/traffic.csv profiles/traffic.json [n] [sample note]
With n given, the sampled column-kernel launches are taken to be columns 0..k-1 of the first tridiagonalisation and
their algorithmic bytes (lower triangle of the trailing matrix once per column) are stored beside the measured ones."""
import csv
import json
import sys
from collections import defaultdict

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
        "ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}
CLASSES = {"gemm_kernel": "gemm_kernel", "gemm_tma_kernel": "gemm_kernel", "trd_symv": "trd_symv_kernel", "trd_dots": "trd_dots_kernel",
           "trd_w_kernel": "trd_w_kernel", "panel_kernel": "panel_kernel"}

with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
agg = defaultdict(lambda: {"ids": set(), "read": 0.0, "write": 0.0, "ms": 0.0})
for r in csv.DictReader(lines):
    name = r["Kernel Name"]
    cls = next((v for k, v in CLASSES.items() if k in name), None)
    if cls is None:
        continue
    a = agg[cls]
    a["ids"].add(r["ID"])
    v = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
    if r["Metric Name"] == "dram__bytes_read.sum":
        a["read"] += v
    elif r["Metric Name"] == "dram__bytes_write.sum":
        a["write"] += v
    elif r["Metric Name"] == "gpu__time_duration.sum":
        a["ms"] += v
out = {"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none; "
                 "python bench.py --steps 1 --warmup 0 --no-cpu (one step = eigh_full!+svd_compact! 8192 f64)",
       "kernels": {}}
for cls, a in agg.items():
    n = len(a["ids"])
    out["kernels"][cls] = {"launches": n, "dram_bytes_per_launch": (a["read"] + a["write"]) / max(n, 1),
                           "dram_read_bytes_total": a["read"], "dram_write_bytes_total": a["write"],
                           "ncu_ms_total": a["ms"]}
n = int(sys.argv[3]) if len(sys.argv) > 3 else 0
note = sys.argv[4] if len(sys.argv) > 4 else ""
for cls, k in out["kernels"].items():
    if note:
        k["sample"] = note
    if cls == "trd_symv_kernel" and n > 0:
        cnt = k["launches"]
        alg = sum(8.0 * (n - c - 1) * (n - c) / 2 for c in range(cnt)) / max(cnt, 1)
        k["alg_bytes_per_launch_same_sample"] = alg
        k["traffic_over_algorithmic"] = k["dram_bytes_per_launch"] / alg
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))
