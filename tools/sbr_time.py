"""Timing of the experimental band -> tridiagonal stage: python tools/sbr_time.py [n] [b ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import makb200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
bs = [int(x) for x in sys.argv[2:]] or [64, 32]
for dtype in (torch.float64, torch.complex128):
    for b in bs:
        g = torch.Generator(device="cuda"); g.manual_seed(1)
        # band matrix in dense storage (only the lower band is read)
        A = torch.randn((n, n), dtype=dtype, device="cuda", generator=g).t()
        ts = []
        for it in range(3):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); d, e, V2, tau2 = makb200.sbr_chase_(A, b); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        # eigenvalue check against the dense band matrix (n <= 4096 only: host eigvalsh)
        msg = ""
        if n <= 4096:
            An = makb200.to_numpy(A)
            i, j = np.indices((n, n))
            B = np.where((i - j >= 0) & (i - j <= b), An, 0)
            B = B + np.tril(B, -1).conj().T
            B[np.diag_indices(n)] = B.diagonal().real
            from scipy.linalg import eigh_tridiagonal
            w = eigh_tridiagonal(d.cpu().numpy(), e.cpu().numpy(), eigvals_only=True)
            wref = np.linalg.eigvalsh(B)
            msg = f" max|dw|/|w|max = {np.abs(w - wref).max() / np.abs(wref).max():.2e}"
        print(f"sbr_chase {str(dtype).split('.')[-1]} n={n} b={b}: {min(ts):.2f} ms (runs {[round(t, 1) for t in ts]}){msg}", flush=True)
