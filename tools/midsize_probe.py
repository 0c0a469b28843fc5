"""Single-matrix svd_compact!/eigh_full! at the mid sizes of config 3 (c128, n = 96..512): time, launches, phases
(MAKB200_PROFILE=1 prints the phase split).  Shows what the pooled per-block path of the batched entry points pays."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import makb200

lib = makb200._lib.load()
for n in (96, 128, 256, 384, 512):
    g = torch.Generator(device="cuda"); g.manual_seed(n)
    A0 = torch.randn((n, n), dtype=torch.complex128, device="cuda", generator=g).t()
    A = makb200.colmajor_empty(n, n, torch.complex128, "cuda")
    USV = makb200.svd.initialize_output(A)
    for it in range(3):
        A.copy_(A0); torch.cuda.synchronize()
        l0 = lib.makb200_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); makb200.svd_compact_(A, USV); e1.record(); torch.cuda.synchronize()
        ms, nl = e0.elapsed_time(e1), lib.makb200_launch_count() - l0
    print(f"svd_compact c128 n={n}: {ms:.3f} ms, {nl} launches ({1e3 * ms / max(nl, 1):.2f} us/launch)", flush=True)
    H0 = (A0 + A0.conj().t()).t().contiguous().t()
    DV = makb200.eigh.initialize_output(A)
    for it in range(3):
        A.copy_(H0); torch.cuda.synchronize()
        l0 = lib.makb200_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); makb200.eigh_full_(A, DV); e1.record(); torch.cuda.synchronize()
        ms, nl = e0.elapsed_time(e1), lib.makb200_launch_count() - l0
    print(f"eigh_full   c128 n={n}: {ms:.3f} ms, {nl} launches ({1e3 * ms / max(nl, 1):.2f} us/launch)", flush=True)
