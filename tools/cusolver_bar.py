"""The GPU bar the reference's own CUDA extension sets: cuSOLVER / cuBLAS library calls (through torch.linalg),
timed on the same box in the same run as our kernels.  Library calls only - never on the product path.
Reference call sites: ext/MatrixAlgebraKitCUDAExt/yacusolver.jl:12-14 (geqrf/ormqr), :158 (syevd), :249/:796 (gesvd/gesvdj/gesvdp).
  python tools/cusolver_bar.py [out.json]      (default gpurun_out/r2_cusolver_bar.json)"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch


def wall(fn, reps=2):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r2_cusolver_bar.json"
    os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
    res = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "unit": "ms", "how": "wall clock around torch.linalg.* with synchronize, best of 2 after 1 warm-up", "rows": []}
    g = torch.Generator(device="cuda")
    g.manual_seed(2)

    def row(name, n, dtype, ms, flops):
        r = {"op": name, "n": n, "dtype": str(dtype).replace("torch.", ""), "ms": round(ms, 3), "tflops": round(flops / ms / 1e9, 2)}
        res["rows"].append(r)
        print(json.dumps(r), flush=True)
        with open(out, "w") as f:
            json.dump(res, f, indent=1)

    for dtype in (torch.float64, torch.complex128):
        c = 4 if dtype == torch.complex128 else 1
        for n in (4096, 8192):
            G = torch.randn((n, n), dtype=dtype, device="cuda", generator=g)
            H = (G + G.conj().t()) / 2
            row("syevd (torch.linalg.eigh)", n, dtype, wall(lambda: torch.linalg.eigh(H)), c * 10 * n ** 3 / 3)
            row("geqrf+orgqr (torch.linalg.qr)", n, dtype, wall(lambda: torch.linalg.qr(G)), c * 8 * n ** 3 / 3)
            row("gemm n^3 (cuBLAS)", n, dtype, wall(lambda: torch.matmul(G, H), reps=3), c * 2 * n ** 3)
            if dtype == torch.float64 or n == 4096:
                for drv in ("gesvdj", "gesvd"):
                    if drv == "gesvd" and n == 8192:
                        continue
                    try:
                        row(f"{drv} (torch.linalg.svd)", n, dtype, wall(lambda: torch.linalg.svd(G, full_matrices=False, driver=drv), reps=1), c * 20 * n ** 3 / 3)
                    except Exception as e:  # noqa: BLE001
                        print("skip", drv, n, repr(e)[:100])
            del G, H
    # the skinny shapes our blocked algorithms actually issue
    for (m, n, k) in [(128, 4096, 4096), (4096, 4096, 128), (8192, 8192, 128), (8192, 8192, 64), (128, 8192, 8192), (1 << 20, 256, 256)]:
        A = torch.randn((m, k), dtype=torch.float64, device="cuda", generator=g)
        B = torch.randn((k, n), dtype=torch.float64, device="cuda", generator=g)
        ms = wall(lambda: torch.matmul(A, B), reps=3)
        r = {"op": f"dgemm {m}x{n}x{k} (cuBLAS)", "ms": round(ms, 4), "tflops": round(2.0 * m * n * k / ms / 1e9, 2)}
        res["rows"].append(r)
        print(json.dumps(r), flush=True)
    with open(out, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
