"""Round-2 bring-up timing of the values-only / leading-rank paths against the full decompositions
(CUDA events, 1 warm-up + 3 timed runs, inputs resident).  usage: vals_time.py <n> <r>"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import makb200  # noqa: E402


def _time(fn, prep, reps=3):
    ms = []
    for i in range(reps + 1):
        x = prep()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn(x)
        b.record()
        torch.cuda.synchronize()
        if i:
            ms.append(a.elapsed_time(b))
    return sorted(ms)[len(ms) // 2]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    r = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    g = torch.Generator(device="cuda").manual_seed(1)
    G = torch.randn((n, n), dtype=torch.float64, device="cuda", generator=g).t()
    H = makb200.colmajor_empty(n, n, torch.float64, "cuda")
    H.copy_((G + G.t()) / 2)

    def cp(src):
        def f():
            x = makb200.colmajor_empty(n, n, torch.float64, "cuda")
            x.copy_(src)
            return x
        return f

    print(f"n={n} f64  eigh_full  {_time(lambda x: makb200.eigh_full_(x), cp(H)):9.1f} ms")
    print(f"n={n} f64  eigh_vals  {_time(lambda x: makb200.eigh_vals_(x), cp(H)):9.1f} ms")
    print(f"n={n} f64  svd_compact {_time(lambda x: makb200.svd_compact_(x), cp(G)):8.1f} ms")
    print(f"n={n} f64  svd_vals   {_time(lambda x: makb200.svd_vals_(x), cp(G)):9.1f} ms")
    print(f"n={n} f64  svd_trunc r={r} (leading) {_time(lambda x: makb200.svd_trunc_(x, None, None, makb200.truncrank(r)), cp(G)):9.1f} ms")
    w = makb200.eigh_vals_(cp(H)())
    w2, _ = makb200.eigh_full_(cp(H)())
    print("max |eigh_vals - eigh_full| / max|w| =", float((w - w2).abs().max() / w2.abs().max()))


if __name__ == "__main__":
    main()
