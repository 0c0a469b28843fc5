"""Small driver for ncu / phase profiling: python tools/prof_run.py {qr|eigh} n [dtype]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import makb200

op, n = sys.argv[1], int(sys.argv[2])
dtype = sys.argv[3] if len(sys.argv) > 3 else "f64"
td = torch.float64 if dtype == "f64" else torch.complex128
reps = int(os.environ.get("REPS", "2"))
G = torch.randn((n, n), dtype=td, device="cuda").t()
if op == "qr":
    A = makb200.colmajor_empty(n, n, td, "cuda")
    Q, R = makb200.qr.initialize_output("qr_compact", A)
    for _ in range(reps):
        A.copy_(G)
        makb200.qr_compact_(A, (Q, R))
elif op == "eigh":
    H = ((G + G.conj().t()) / 2).t().contiguous().t()
    A = makb200.colmajor_empty(n, n, td, "cuda")
    D, V = makb200.eigh.initialize_output(A)
    for _ in range(reps):
        A.copy_(H)
        makb200.eigh_full_(A, (D, V))
torch.cuda.synchronize()
if os.environ.get("KTIME"):
    import ctypes
    lib = makb200._lib.load()
    lib.makb200_kernel_timing(1)
    if op == "eigh":
        A.copy_(H); makb200.eigh_full_(A, (D, V))
    else:
        A.copy_(G); makb200.qr_compact_(A, (Q, R))
    torch.cuda.synchronize()
    for which, nm in ((0, "symv"), (1, "gemm"), (2, "w")):
        ms, nl = ctypes.c_double(), ctypes.c_int()
        lib.makb200_kernel_time(which, ctypes.byref(ms), ctypes.byref(nl))
        print(f"[ktime] {nm}: {ms.value:.2f} ms over {nl.value} launches, gemm_flops={lib.makb200_gemm_flops():.3e}")
    lib.makb200_kernel_timing(0)
print("done")
