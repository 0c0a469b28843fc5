"""Measure the FP64 tensor-core denominator on this box: cuBLAS DGEMM 8192^3 (burst: best of 10, and
sustained: back to back for ~3 s), and a STREAM-style copy for reference. Writes profiles/fp64_peak.json.
cuBLAS is used here as a yardstick only — never on the product path."""
import json, os, sys, time
import torch
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
c = torch.empty_like(a)
for _ in range(3): torch.matmul(a, b, out=c)
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 100
e0.record()
for _ in range(reps): torch.matmul(a, b, out=c)
e1.record(); torch.cuda.synchronize()
sus = e0.elapsed_time(e1) / reps
out = {"dgemm_tflops": 2 * n ** 3 / best / 1e9, "dgemm_tflops_sustained": 2 * n ** 3 / sus / 1e9,
       "how": "torch.matmul float64 8192^3 (cuBLAS), best of 10 / 100 back to back, CUDA events",
       "gpu": torch.cuda.get_device_name(0)}
os.makedirs("profiles", exist_ok=True)
json.dump(out, open("gpurun_out/fp64_peak.json", "w"), indent=1)
print(json.dumps(out))
