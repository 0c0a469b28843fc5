"""One batched call on one size bucket (for ncu launch lists / single-kernel captures).
  python tools/batched_prof.py qr 257 512 [nblocks] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import makb200

op, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
nblocks = int(sys.argv[4]) if len(sys.argv) > 4 else 20000
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 1
dev = torch.device("cuda", 0)
rng = np.random.Generator(np.random.PCG64(4))
dims = np.rint(16 * 32 ** rng.random(nblocks)).astype(int)
dims = [int(n) for n in dims if lo <= n <= hi]
g = torch.Generator(device=dev); g.manual_seed(4)
blocks = [torch.randn((n, n), dtype=torch.complex128, device=dev, generator=g).t() for n in dims]
if op == "eigh":
    blocks = [a + a.conj().t() for a in blocks]
As = [makb200.colmajor_empty(a.shape[0], a.shape[1], a.dtype, dev) for a in blocks]
plan = {"qr": makb200.BatchedQRPlan, "svd": makb200.BatchedSVDPlan, "eigh": makb200.BatchedEighPlan}[op](As)
for it in range(reps):
    for a, b in zip(As, blocks):
        a.copy_(b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); plan.run(); e1.record(); torch.cuda.synchronize()
    print(f"{op} {lo}-{hi}: {len(dims)} blocks, {e0.elapsed_time(e1):.3f} ms", flush=True)
