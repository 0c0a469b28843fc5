"""Quick device-side timing of the building blocks (CUDA events). Not the bench contract."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import makb200
from oracle import mak_oracle as O


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def randdev(m, n, dtype):
    t = torch.randn((n, m), dtype=torch.float64 if dtype == "f64" else torch.complex128, device="cuda")
    return t.t()


def main():
    print(torch.cuda.get_device_name(0))
    which = sys.argv[1:] or ["gemm", "qr"]
    if "gemm" in which:
        for dtype in ("f64", "c128"):
            for (m, n, k) in [(4096, 4096, 4096), (8192, 8192, 8192), (4096, 4096, 128), (128, 4096, 4096), (8192, 8192, 128)]:
                if dtype == "c128" and m * n > 4096 * 4096 + 1 and k > 128: continue
                A, B, C = randdev(m, k, dtype), randdev(k, n, dtype), randdev(m, n, dtype)
                ms = timeit(lambda: makb200.gemm_(C, A, B, 1.0, 0.0))
                fl = 2.0 * m * n * k * (4 if dtype == "c128" else 1)
                tms = timeit(lambda: torch.matmul(A, B))
                print(f"gemm {dtype} {m}x{n}x{k}: {ms:.3f} ms  {fl/ms/1e9:.1f} TF/s   (cuBLAS via torch: {tms:.3f} ms {fl/tms/1e9:.1f} TF/s)")
            A, B, C = randdev(4096, 4096, dtype), randdev(4096, 4096, dtype), randdev(4096, 4096, dtype)
            ms = timeit(lambda: makb200.gemm_(C, A, B, 1.0, 0.0, "C", "N"))
            print(f"gemm {dtype} C,N 4096^3: {ms:.3f} ms {2*4096**3*(4 if dtype=='c128' else 1)/ms/1e9:.1f} TF/s")
    if "qr" in which:
        for dtype in ("f64", "c128"):
            for (m, n) in [(1024, 1024), (4096, 4096), (8192, 8192), (16384, 256)]:
                if dtype == "c128" and m == 8192: continue
                A0 = randdev(m, n, dtype)
                A = makb200.colmajor_empty(m, n, A0.dtype, A0.device)
                Q, R = makb200.qr.initialize_output("qr_compact", A)
                def run():
                    A.copy_(A0)
                    makb200.qr_compact_(A, (Q, R))
                def cp():
                    A.copy_(A0)
                ms = timeit(run) - timeit(cp)
                c = 4 if dtype == "c128" else 1
                fl = c * (4.0 * m * n * n - 4.0 * n ** 3 / 3)
                print(f"qr_compact {dtype} {m}x{n}: {ms:.3f} ms  {fl/ms/1e9:.2f} TF/s")
                if m <= 4096:
                    An, Qn, Rn = makb200.to_numpy(A0), makb200.to_numpy(Q), makb200.to_numpy(R)
                    print("   resid", O.rel_resid(An, Qn, Rn), "orth", O.orth_err(Qn), "tol", O.tol_for(m, n))
                    def tq():
                        torch.linalg.qr(A0)
                    print(f"   torch.linalg.qr (cuSOLVER): {timeit(tq, reps=3, warm=1):.3f} ms")




def eigh_probe():
    import time
    for dtype in ("f64", "c128"):
        for n in (1024, 4096, 8192):
            if dtype == "c128" and n == 8192 and "big" not in sys.argv: continue
            G = randdev(n, n, dtype)
            A0 = ((G + G.conj().t()) / 2).t().contiguous().t()
            A = makb200.colmajor_empty(n, n, A0.dtype, A0.device)
            D, V = makb200.eigh.initialize_output(A)
            def run():
                A.copy_(A0)
                makb200.eigh_full_(A, (D, V))
            def cp():
                A.copy_(A0)
            ms = timeit(run, reps=3, warm=1) - timeit(cp, reps=3, warm=1)
            c = 4 if dtype == "c128" else 1
            fl = c * 10.0 * n ** 3 / 3
            print(f"eigh_full {dtype} n={n}: {ms:.2f} ms  {fl/ms/1e9:.2f} TF/s", flush=True)
            w = D.cpu().numpy()
            if n <= 4096:
                An, Vn = makb200.to_numpy(A0), makb200.to_numpy(V)
                print("   resid", np.linalg.norm(An @ Vn - Vn * w) / np.linalg.norm(An), "orth", O.orth_err(Vn), "tol", O.tol_for(n))
            t0 = time.time(); torch.linalg.eigh(A0); torch.cuda.synchronize(); t1 = time.time()
            torch.linalg.eigh(A0); torch.cuda.synchronize(); t2 = time.time()
            print(f"   torch.linalg.eigh (cuSOLVER syevd): {(t2-t1)*1e3:.1f} ms")


if "eigh" in sys.argv[1:]:
    eigh_probe()

if __name__ == "__main__":
    if any(a in ("gemm", "qr") for a in sys.argv[1:]) or len(sys.argv) == 1:
        main()


def svd_probe():
    for dtype in ("f64",):
        for n in (2048, 4096, 8192):
            A0 = randdev(n, n, dtype)
            A = makb200.colmajor_empty(n, n, A0.dtype, A0.device)
            USV = makb200.svd.initialize_output(A)
            def run():
                A.copy_(A0)
                makb200.svd_compact_(A, USV)
            ms = timeit(run, reps=2, warm=1)
            fl = 20.0 * n ** 3 / 3
            print(f"svd_compact {dtype} n={n}: {ms:.2f} ms  {fl/ms/1e9:.2f} TF/s", flush=True)
            WP = makb200.polar.initialize_output(A)
            def runp():
                A.copy_(A0)
                makb200.left_polar_(A, WP)
            ms = timeit(runp, reps=2, warm=1)
            print(f"left_polar {dtype} n={n}: {ms:.2f} ms", flush=True)
            if n <= 4096:
                import time
                torch.linalg.svd(A0); torch.cuda.synchronize(); t1 = time.time()
                torch.linalg.svd(A0); torch.cuda.synchronize(); t2 = time.time()
                print(f"   torch.linalg.svd (cuSOLVER gesvdj/gesvd): {(t2-t1)*1e3:.1f} ms")


if "svd" in sys.argv[1:]:
    svd_probe()
