"""Scale invariance probe: svd_compact / eigh_full / left_polar (single and batched) on matrices scaled by 1e-100 .. 1e100.
Prints the error of every path relative to the scaled norm."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import makb200
from oracle import mak_oracle as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 111
for dtype in ("f64", "c128"):
    for sc in (1e-100, 1e-30, 1.0, 1e30, 1e100):
        A = np.asfortranarray(sc * O.randn_matrix(n, n, dtype, 7))
        H = np.asfortranarray(sc * O.rand_hermitian(n, dtype, seed=8))
        so = np.linalg.svd(A, compute_uv=False)
        U, S, Vh = makb200.svd_compact(makb200.to_device(A))
        Sn = S.cpu().numpy()
        r1 = np.max(np.abs(Sn - so)) / so[0]
        r2 = np.linalg.norm(A - (makb200.to_numpy(U) * Sn) @ makb200.to_numpy(Vh)) / np.linalg.norm(A)
        D, V = makb200.eigh_full(makb200.to_device(H))
        w = D.cpu().numpy()
        r3 = np.max(np.abs(w - np.linalg.eigvalsh(H))) / np.abs(w).max()
        W, P = makb200.left_polar(makb200.to_device(A))
        r4 = np.linalg.norm(makb200.to_numpy(W) @ makb200.to_numpy(P) - A) / np.linalg.norm(A)
        res = []
        for ls in ("1", "0"):
            os.environ["MAKB200_SVD_LOCKSTEP"] = ls
            outs = makb200.svd_compact_batched_([makb200.to_device(A) for _ in range(9)])
            Sb = outs[3][1].cpu().numpy()
            res.append(np.max(np.abs(Sb - so)) / so[0])
        outs = makb200.eigh_full_batched_([makb200.to_device(H) for _ in range(9)])
        r7 = np.max(np.abs(outs[2][0].cpu().numpy() - np.linalg.eigvalsh(H))) / np.abs(w).max()
        print(f"{dtype} scale {sc:.0e}: svd dS {r1:.2e} resid {r2:.2e} | eigh dw {r3:.2e} | polar resid {r4:.2e} | "
              f"batched svd dS lockstep {res[0]:.2e} per-block {res[1]:.2e} | batched eigh dw {r7:.2e}", flush=True)
