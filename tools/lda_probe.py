"""eigh_full! 8192 f64 with different leading dimensions of A (HBM channel mapping of the 16-column tiles)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import makb200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
g = torch.Generator(device="cuda"); g.manual_seed(2)
G = torch.randn((n, n), dtype=torch.float64, device="cuda", generator=g)
A0 = (G + G.t()) / 2
del G
for pad in (0, 8, 16, 32, 64, 136):
    big = makb200.colmajor_empty(n + pad, n, torch.float64, "cuda")
    A = big[:n, :]
    D = torch.empty(n, dtype=torch.float64, device="cuda")
    V = makb200.colmajor_empty(n, n, torch.float64, "cuda")
    ts = []
    for it in range(3):
        A.copy_(A0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); makb200.eigh_full_(A, (D, V)); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    w = D
    res = float(torch.linalg.matrix_norm(A0 @ V - V * w) / torch.linalg.matrix_norm(A0))
    print(f"n={n} lda={n + pad}: eigh_full {min(ts):.1f} ms  resid {res:.1e}", flush=True)
    del big, A, V
