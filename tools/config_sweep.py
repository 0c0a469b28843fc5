"""Every single-GPU BASELINE config at its FULL size, with size-independent parity properties
(residual, orthogonality, ordering, gauge) evaluated on the device, and device timings.
  python tools/config_sweep.py [C1] [C2] [C2c] [C5]      (default: all)
Prints one JSON line per case; not the bench contract (bench.py is)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import makb200

EPS = 2.220446049250313e-16
dev = torch.device("cuda", 0)


def gauss(m, n, dtype, seed):
    g = torch.Generator(device=dev); g.manual_seed(seed)
    t = torch.randn((n, m), dtype=torch.float64 if dtype == "f64" else torch.complex128, device=dev, generator=g)
    return t.t()        # column-major m x n


def timed(fn, reps):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), out


def fro(x):
    return float(torch.linalg.matrix_norm(x).item())


def orth(Q, side="left"):
    G = Q.conj().t() @ Q if side == "left" else Q @ Q.conj().t()
    G.diagonal().sub_(1.0)
    return fro(G)


def emit(**kw):
    print(json.dumps(kw), flush=True)


def case_qr(m, n, dtype, seed, name, reps=5):
    A0 = gauss(m, n, dtype, seed)
    A = makb200.colmajor_empty(m, n, A0.dtype, dev)
    QR = makb200.qr.initialize_output("qr_compact", A)
    def run():
        A.copy_(A0)
        return makb200.qr_compact_(A, QR)
    run()
    ms, (Q, R) = timed(run, reps)
    cp, _ = timed(lambda: A.copy_(A0), reps)
    ms -= cp
    c = 4 if dtype == "c128" else 1
    fl = c * (4.0 * m * n * n - 4.0 * n ** 3 / 3)
    tol = 10 * max(m, n) * EPS
    d = torch.diagonal(R)
    emit(case=name, op="qr_compact!", dtype=dtype, m=m, n=n, ms=ms, TFLOPs=fl / ms / 1e9,
         resid=fro(A0 - Q @ R) / fro(A0), orth=orth(Q), tril=fro(torch.tril(R, -1)),
         diag_nonneg=bool((d.real >= 0).all().item()), tol=tol)


def case_eigh(n, dtype, seed, name, reps=3):
    G = gauss(n, n, dtype, seed)
    A0 = ((G + G.conj().t()) / 2).t().contiguous().t()
    del G
    A = makb200.colmajor_empty(n, n, A0.dtype, dev)
    DV = makb200.eigh.initialize_output(A)
    def run():
        A.copy_(A0)
        return makb200.eigh_full_(A, DV)
    run()
    ms, (D, V) = timed(run, reps)
    c = 4 if dtype == "c128" else 1
    fl = c * 10.0 * n ** 3 / 3
    w = torch.diagonal(D) if D.dim() == 2 else D
    w = w.to(V.dtype)
    emit(case=name, op="eigh_full!", dtype=dtype, n=n, ms=ms, TFLOPs=fl / ms / 1e9,
         resid=fro(A0 @ V - V * w) / fro(A0), orth=orth(V), ascending=bool((torch.diff(w.real) >= 0).all().item()),
         tol=10 * n * EPS)


def case_svd(n, dtype, seed, name, reps=2, trunc=None):
    A0 = gauss(n, n, dtype, seed)
    A = makb200.colmajor_empty(n, n, A0.dtype, dev)
    USV = makb200.svd.initialize_output(A)
    def run():
        A.copy_(A0)
        return makb200.svd_compact_(A, USV)
    run()
    ms, (U, S, Vh) = timed(run, reps)
    c = 4 if dtype == "c128" else 1
    fl = c * 20.0 * n ** 3 / 3
    s = (torch.diagonal(S) if S.dim() == 2 else S)
    emit(case=name, op="svd_compact!", dtype=dtype, n=n, ms=ms, TFLOPs=fl / ms / 1e9,
         resid=fro(A0 - (U * s.to(U.dtype)) @ Vh) / fro(A0), orth_U=orth(U), orth_V=orth(Vh, "right"),
         descending=bool((torch.diff(s) <= 0).all().item()), smin=float(s.min().item()), tol=10 * n * EPS)
    if trunc:
        s_full = s.clone()
        def runt():
            A.copy_(A0)
            return makb200.svd_trunc(A, trunc=makb200.truncrank(trunc))
        runt()
        ms, out = timed(runt, reps)
        Ut, St, Vht, eps_t = out
        st = (torch.diagonal(St) if St.dim() == 2 else St)
        tail = float(torch.linalg.vector_norm(s_full[trunc:]).item())
        emit(case=name, op=f"svd_trunc!(truncrank({trunc}))", dtype=dtype, n=n, ms=ms, TFLOPs=fl / ms / 1e9,
             shapes=[list(Ut.shape), list(st.shape), list(Vht.shape)],
             vals_match=float((st - s_full[:trunc]).abs().max().item() / s_full[0].item()),
             eps=float(eps_t), eps_expected=tail,
             resid_trunc=abs(fro(A0 - (Ut * st.to(Ut.dtype)) @ Vht) - tail) / fro(A0), tol=10 * n * EPS)


def case_polar(n, dtype, seed, name, reps=2):
    A0 = gauss(n, n, dtype, seed)
    A = makb200.colmajor_empty(n, n, A0.dtype, dev)
    WP = makb200.polar.initialize_output(A)
    def run():
        A.copy_(A0)
        return makb200.left_polar_(A, WP)
    run()
    ms, (W, P) = timed(run, reps)
    emit(case=name, op="left_polar! (QDWH)", dtype=dtype, n=n, ms=ms,
         resid=fro(W @ P - A0) / fro(A0), orth=orth(W), herm=fro(P - P.conj().t()) / fro(P), tol=10 * n * EPS)


which = set(sys.argv[1:]) or {"C1", "C2", "C2c", "C5"}
print(torch.cuda.get_device_name(0), flush=True)
if "C1" in which:
    case_qr(4096, 4096, "f64", 1, "C1")
if "C2" in which:
    case_eigh(8192, "f64", 2, "C2")
    case_svd(8192, "f64", 2, "C2")
if "C2c" in which:
    case_eigh(8192, "c128", 3, "C2")
    case_svd(8192, "c128", 3, "C2")
if "C5" in which:
    case_polar(16384, "f64", 6, "C5")
    case_svd(16384, "f64", 6, "C5", reps=1, trunc=1024)
