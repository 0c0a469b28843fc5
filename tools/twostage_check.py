"""Bring-up of the EXPERIMENTAL two-stage eigh path, stage by stage (diagnostic prints, no asserts):
  1. sy2sb_: eigenvalues of the resulting band matrix vs the dense matrix
  2. sbr_apply_q2_: GPU diamond-blocked Q2 Z vs a host reflector loop
  3. eigh_full with MAKB200_EIGH_TWOSTAGE set (the env var is read once per process: set it on the
     command line): residual / orthogonality / eigenvalue error, then timings with phase breakdown.
  MAKB200_EIGH_TWOSTAGE=64 python tools/twostage_check.py [nbig]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import makb200

EPS = np.finfo(float).eps
b = int(os.environ.get("MAKB200_EIGH_TWOSTAGE", "64"))
b = 64 if b == 1 else max(8, min(64, b))


def herm(n, dtype, seed):
    rng = np.random.default_rng(seed)
    G = rng.standard_normal((n, n))
    if dtype == "c128":
        G = G + 1j * rng.standard_normal((n, n))
    return np.asfortranarray((G + G.conj().T) / 2)


def host_q2(V2, tau2, n, b, Z):
    X = Z.astype(V2.dtype).copy()
    for s in range(n - 2, -1, -1):
        for k in range((n - 1 - s + b - 1) // b - 1, -1, -1):
            r0 = s + 1 + k * b
            L = min(b, n - r0)
            v, tau = V2[r0:r0 + L, s], tau2[k, s]
            X[r0:r0 + L] -= np.outer(v, tau * (v.conj() @ X[r0:r0 + L]))
    return X


for dtype in (() if os.environ.get("SKIP_SMALL") else ("f64", "c128")):
    for n in (150, 333):
        A0 = herm(n, dtype, n)
        w0 = np.linalg.eigvalsh(A0)
        # 1. stage 1
        A = makb200.to_device(A0)
        makb200.sy2sb_(A, b)
        torch.cuda.synchronize()
        An = makb200.to_numpy(A)
        i, j = np.indices((n, n))
        B = np.where((i - j >= 0) & (i - j <= b), An, 0)
        B = B + np.tril(B, -1).conj().T
        B[np.diag_indices(n)] = B.diagonal().real
        print(f"[{dtype} n={n} b={b}] sy2sb: max|eig(B)-eig(A)|/|w| = {np.abs(np.linalg.eigvalsh(B) - w0).max() / np.abs(w0).max():.2e}", flush=True)
        # 2. chase + Q2 on the GPU vs host loop
        Bd = makb200.to_device(np.asfortranarray(B))
        d, e, V2, tau2 = makb200.sbr_chase_(Bd, b)
        rng = np.random.default_rng(1)
        Z0 = rng.standard_normal((n, 40)) + (1j * rng.standard_normal((n, 40)) if dtype == "c128" else 0)
        Zd = makb200.to_device(np.asfortranarray(Z0))
        makb200.sbr_apply_q2_(V2, tau2, b, Zd)
        torch.cuda.synchronize()
        Xh = host_q2(makb200.to_numpy(V2), makb200.to_numpy(tau2), n, b, Z0)
        print(f"[{dtype} n={n} b={b}] apply_q2: ||gpu - host|| / ||host|| = {np.linalg.norm(makb200.to_numpy(Zd) - Xh) / np.linalg.norm(Xh):.2e}", flush=True)
        # 3. assembled path (only two-stage if the env var is set and n > 2b)
        D, V = makb200.eigh_full(makb200.to_device(A0))
        torch.cuda.synchronize()
        w = (torch.diagonal(D) if D.dim() == 2 else D).cpu().numpy().real
        Vn = makb200.to_numpy(V)
        print(f"[{dtype} n={n}] eigh_full (two-stage={'MAKB200_EIGH_TWOSTAGE' in os.environ and n > 2 * b}): "
              f"vals {np.abs(w - w0).max() / np.abs(w0).max():.2e} resid {np.linalg.norm(A0 @ Vn - Vn * w) / np.abs(w0).max():.2e} "
              f"orth {np.linalg.norm(Vn.conj().T @ Vn - np.eye(n)):.2e} tol {10 * n * EPS:.2e}", flush=True)

nbig = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
for dtype in (torch.float64,):
    g = torch.Generator(device="cuda"); g.manual_seed(2)
    G = torch.randn((nbig, nbig), dtype=dtype, device="cuda", generator=g)
    A0 = ((G + G.conj().t()) / 2).t().contiguous().t()
    A = makb200.colmajor_empty(nbig, nbig, dtype, "cuda")
    DV = makb200.eigh.initialize_output(A)
    for it in range(2):
        A.copy_(A0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); D, V = makb200.eigh_full_(A, DV); e1.record(); torch.cuda.synchronize()
    w = (torch.diagonal(D) if D.dim() == 2 else D).to(V.dtype)
    res = float(torch.linalg.matrix_norm(A0 @ V - V * w) / torch.linalg.matrix_norm(A0))
    Gm = V.conj().t() @ V
    Gm.diagonal().sub_(1.0)
    print(f"eigh_full n={nbig}: {e0.elapsed_time(e1):.1f} ms resid {res:.2e} orth {float(torch.linalg.matrix_norm(Gm)):.2e}", flush=True)
