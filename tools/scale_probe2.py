"""Narrowing the f64 extreme-scale failure of the phased batched SVD."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import makb200
from oracle import mak_oracle as O

np.set_printoptions(precision=3, linewidth=200)
for n in (111, 112, 200):
    for sc in (1e-100, 1e-60, 1e-40):
        A = np.asfortranarray(sc * O.randn_matrix(n, n, "f64", 7))
        H = np.asfortranarray(sc * O.rand_hermitian(n, "f64", seed=8))
        so = np.linalg.svd(A, compute_uv=False)
        line = f"n={n} scale {sc:.0e}:"
        for name, env in (("lockstep", {"MAKB200_SVD_LOCKSTEP": "1"}), ("phased", {"MAKB200_SVD_LOCKSTEP": "0"}),
                          ("pooled", {"MAKB200_SVD_PHASED": "0"}), ("nobhetrd", {"MAKB200_SVD_PHASED": "0", "MAKB200_BHETRD": "0"})):
            for k in ("MAKB200_SVD_LOCKSTEP", "MAKB200_SVD_PHASED", "MAKB200_BHETRD"):
                os.environ.pop(k, None)
            os.environ.update(env)
            outs = makb200.svd_compact_batched_([makb200.to_device(A) for _ in range(9)])
            Sb = outs[3][1].cpu().numpy()
            line += f" {name} {np.max(np.abs(Sb - so)) / so[0]:.2e}"
            if name == "phased" and sc == 1e-100 and n == 111:
                print("S/so head", (Sb / so)[:6], "tail", (Sb / so)[-4:])
        for k in ("MAKB200_SVD_LOCKSTEP", "MAKB200_SVD_PHASED", "MAKB200_BHETRD"):
            os.environ.pop(k, None)
        outs = makb200.eigh_full_batched_([makb200.to_device(H) for _ in range(9)])
        wr = np.linalg.eigvalsh(H)
        line += f" | batched eigh {np.max(np.abs(outs[2][0].cpu().numpy() - wr)) / np.abs(wr).max():.2e}"
        print(line, flush=True)
