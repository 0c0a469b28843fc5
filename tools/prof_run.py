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
print("done")
