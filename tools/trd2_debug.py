"""Small-n sweep of eigh_full (values vs numpy) - run once with the persistent kernels and once with MAKB200_SYMV_V2=0."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import makb200
from oracle import mak_oracle as O

bad = 0
for dtype in ("f64", "c128"):
    for n in list(range(2, 40)) + [63, 64, 65, 66, 127, 128, 129, 130, 255, 256, 257, 258, 300, 511, 512, 513, 514, 600, 1023, 1024, 1026]:
        A0 = O.rand_hermitian(n, dtype, seed=123 + n)
        D, V = makb200.eigh_full(makb200.to_device(A0))
        torch.cuda.synchronize()
        w = D.cpu().numpy(); Vn = makb200.to_numpy(V)
        wo = np.linalg.eigvalsh(A0)
        err = np.abs(w - wo).max() / np.abs(wo).max()
        res = np.linalg.norm(A0 @ Vn - Vn * w) / np.linalg.norm(A0)
        ok = err < 1e-12 and res < 1e-12
        if not ok:
            bad += 1
            print(f"BAD {dtype} n={n}: val err {err:.2e} resid {res:.2e}", flush=True)
print("v2 =", os.environ.get("MAKB200_SYMV_V2", "1"), "bad cases:", bad)
