"""Small batched eigh + svd through the lock-step paths (for compute-sanitizer runs)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import makb200
from oracle import mak_oracle as O

os.environ["MAKB200_LOCKSTEP_VERBOSE"] = "1"
for dtype in ("f64", "c128"):
    ns = [131, 140, 150, 133, 170, 129, 200, 137, 145]
    Hs = [O.rand_hermitian(n, dtype, seed=i) for i, n in enumerate(ns)]
    outs = makb200.eigh_full_batched_([makb200.to_device(a) for a in Hs], check=False)
    torch.cuda.synchronize()
    e = max(np.max(np.abs(D.cpu().numpy() - np.linalg.eigvalsh(a))) for a, (D, V) in zip(Hs, outs))
    r = max(np.linalg.norm(a @ makb200.to_numpy(V) - makb200.to_numpy(V) * D.cpu().numpy()) for a, (D, V) in zip(Hs, outs))
    print(dtype, "eigh max dw", e, "max resid", r, flush=True)
    As = [O.randn_matrix(n + (7 if i % 3 == 0 else 0), n, dtype, seed=50 + i) for i, n in enumerate(ns)]
    outs = makb200.svd_compact_batched_([makb200.to_device(a) for a in As])
    torch.cuda.synchronize()
    e = max(np.max(np.abs(S.cpu().numpy() - np.linalg.svd(a, compute_uv=False))) for a, (U, S, Vh) in zip(As, outs))
    r = max(np.linalg.norm(a - (makb200.to_numpy(U) * S.cpu().numpy()) @ makb200.to_numpy(Vh)) for a, (U, S, Vh) in zip(As, outs))
    print(dtype, "svd max dS", e, "max resid", r, flush=True)
