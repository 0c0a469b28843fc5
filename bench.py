#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 dense-factorization hot path.

Contract:
  python bench.py --gpus N --steps K --warmup W          -> one JSON line (our arm)
  python bench.py --impl reference --gpus N --steps K ... -> same line for the reference's CPU path

Workload at N=1 = BASELINE.json configs[1]: `eigh_full!` (and `svd_compact!` when --ops includes
it) on an 8192 x 8192 Float64 matrix.  A "step" is one pass of the hot path over one synthetic
matrix (i.i.d. N(0,1), Hermitian part for eigh).  `value` = algorithmic GFLOP/s
(F_eigh_full = 10 n^3/3, F_svd_compact = 20 n^3/3) with the input resident in HBM; `e2e` = the
same through the public operator with HOST buffers (pinned H2D of A and D2H of all outputs inside
the timed region).  N>1: the configuration does not shard (SURVEY §8e "replicas only"), every rank
factorizes its own matrix; value = aggregate.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "GFLOP/s (qr/svd/eigh Float64 factorizations, algorithmic flops)"


def flops(op, n, cplx=False):
    c = 4.0 if cplx else 1.0
    if op == "eigh":
        return c * 10.0 * n ** 3 / 3.0
    if op == "svd":
        return c * 20.0 * n ** 3 / 3.0
    if op == "qr":
        return c * 8.0 * n ** 3 / 3.0
    raise ValueError(op)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def fp64_peak():
    """FP64 tensor (DMMA) denominator: MEASURED_PEAKS.json carries no FP64 figure, so the measured
    cuBLAS DGEMM 8192^3 rate on this pool's B200 (profiles/fp64_peak.json, tools/peak_fp64.py) is
    used; fallback = the 36.2 TF/s measured in round 1."""
    p = os.path.join(ROOT, "profiles", "fp64_peak.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["dgemm_tflops"]), "measured cuBLAS DGEMM 8192^3 (profiles/fp64_peak.json)"
    return 36.2, "round-1 measurement of cuBLAS DGEMM 8192^3 (no FP64 entry in MEASURED_PEAKS.json)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------
# reference arm: the oracle (LAPACK replay of the reference's call sequence) on host cores
# --------------------------------------------------------------------------------------
def cpu_sample(ops, n_cpu, reps=1):
    """time the oracle on a bounded sample; returns (GFLOP/s, seconds, description)."""
    from oracle import mak_oracle as O
    tot_f, tot_t = 0.0, 0.0
    for op in ops:
        if op == "eigh":
            A = O.rand_hermitian(n_cpu, "f64", seed=2)
            fn = lambda: O.eigh_full(A)  # noqa: E731
        elif op == "svd":
            A = O.randn_matrix(n_cpu, n_cpu, "f64", seed=2)
            fn = lambda: O.svd_compact(A)  # noqa: E731
        else:
            A = O.randn_matrix(n_cpu, n_cpu, "f64", seed=1)
            fn = lambda: O.qr_compact(A)  # noqa: E731
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            tot_t += time.perf_counter() - t0
            tot_f += flops(op, n_cpu)
    return tot_f / tot_t / 1e9, tot_t, f"{'+'.join(ops)} {n_cpu}x{n_cpu} f64 x{reps} (LAPACK via scipy/OpenBLAS)"


def threads_info():
    try:
        from threadpoolctl import threadpool_info
        for lib in threadpool_info():
            if lib.get("user_api") == "blas":
                return int(lib.get("num_threads", os.cpu_count() or 1))
    except Exception:
        pass
    return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ops = args.ops.split(",")
    n_cpu = args.cpu_n
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_sample(ops, min(n_cpu, 1024))
    vals, secs = [], []
    for _ in range(args.steps):
        g, s, desc = cpu_sample(ops, n_cpu)
        vals.append(g); secs.append(s)
    v = float(np.sum([flops(o, n_cpu) for o in ops]) * len(vals) / np.sum(secs) / 1e9)
    cores = threads_info()
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "GFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(secs) * 1e3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{'+'.join(ops)}_full n={args.n} f64 (BASELINE configs[1]); reference step = bounded "
                               f"sample n={n_cpu}, GFLOP/s is size-normalised"},
        "cpu_baseline": {"value": v, "unit": "GFLOP/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": v, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
def run_ours(args):
    import ctypes
    import torch
    import torch.distributed as dist
    import makb200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ops = args.ops.split(",")
    n = args.n
    lib = makb200._lib.load()

    # synthetic inputs: i.i.d. N(0,1), numpy PCG64, seed 2+rank; eigh input = (G+G^T)/2
    # (generated here: the product arm does not touch oracle/)
    host_in = {}
    for op in ops:
        G = np.asfortranarray(np.random.Generator(np.random.PCG64(2 + rank)).standard_normal((n, n)).T)
        host_in[op] = np.asfortranarray((G + G.T) / 2) if op == "eigh" else G
    pinned = {op: torch.from_numpy(np.ascontiguousarray(a.T)).pin_memory() for op, a in host_in.items()}
    dev_in = {op: t.to(dev).t() for op, t in pinned.items()}
    A = makb200.colmajor_empty(n, n, torch.float64, dev)
    outs = {}
    if "eigh" in ops:
        outs["eigh"] = makb200.eigh.initialize_output(A)
    if "svd" in ops:
        outs["svd"] = makb200.svd.initialize_output(A)
    if "qr" in ops:
        outs["qr"] = makb200.qr.initialize_output("qr_compact", A)

    def step_dev():
        for op in ops:
            A.copy_(dev_in[op])  # the op destroys A: refresh from the resident copy (0.2 ms, counted)
            if op == "eigh":
                # check=False would skip the Hermitian pre-check; the reference runs it, so do we
                makb200.eigh_full_(A, outs[op])
            elif op == "svd":
                makb200.svd_compact_(A, outs[op])
            else:
                makb200.qr_compact_(A, outs[op])

    host_out = {}
    for op in ops:
        host_out[op] = [torch.empty(tuple(reversed(o.shape)) if o.dim() == 2 else o.shape, dtype=o.dtype).pin_memory()
                        for o in outs[op]]

    def step_e2e():
        h2d = d2h = 0
        for op in ops:
            A.t().copy_(pinned[op], non_blocking=True)
            h2d += pinned[op].numel() * 8
            if op == "eigh":
                makb200.eigh_full_(A, outs[op])
            elif op == "svd":
                makb200.svd_compact_(A, outs[op])
            else:
                makb200.qr_compact_(A, outs[op])
            for o, ho in zip(outs[op], host_out[op]):
                (ho.copy_(o.t(), non_blocking=True) if o.dim() == 2 else ho.copy_(o, non_blocking=True))
                d2h += o.numel() * o.element_size()
        return h2d, d2h

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    nwarm = max(args.warmup, 3) if not os.environ.get("MAKB200_BENCH_UNDER_NCU") else args.warmup
    for _ in range(nwarm):
        step_dev()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = lib.makb200_launch_count()
    ms = timed(step_dev, args.steps)
    launches = int(lib.makb200_launch_count() - l0)
    clocks = sampler.stop() if rank == 0 else None
    step_flops = sum(flops(op, n) for op in ops)
    value = world * step_flops * args.steps / (ms * 1e-3) / 1e9

    # e2e through the public operators with host buffers
    step_e2e()
    bytes_io = [0, 0]

    def e2e_fn():
        a, b = step_e2e()
        bytes_io[0], bytes_io[1] = a, b
    e2e_steps = max(1, min(args.steps, 3))
    ms_e2e = timed(e2e_fn, e2e_steps)
    e2e_value = world * step_flops * e2e_steps / (ms_e2e * 1e-3) / 1e9

    # roofline of the dominant kernel (tridiagonalisation column-dot kernel, HBM-bound): one extra
    # step with CUDA events around every launch of that kernel (not part of `value`)
    # roofline of the dominant kernel: one extra step with CUDA events around every launch of the two
    # candidate kernels (not part of `value`): the tridiagonalisation column-dot kernel (HBM-bound)
    # and the DMMA GEMM (FP64 tensor-bound)
    lib.makb200_kernel_timing(1)
    step_dev()
    torch.cuda.synchronize()
    kms, kl = ctypes.c_double(), ctypes.c_int()
    lib.makb200_kernel_time(0, ctypes.byref(kms), ctypes.byref(kl))
    gms, gl = ctypes.c_double(), ctypes.c_int()
    lib.makb200_kernel_time(1, ctypes.byref(gms), ctypes.byref(gl))
    gflops = float(lib.makb200_gemm_flops())
    lib.makb200_kernel_timing(0)
    step_ms = ms / args.steps
    hbm_peak, how = peaks()
    cands = []
    if kl.value > 0:
        # algorithmic bytes: every DISTINCT element of the Hermitian trailing matrix once per column step,
        # for each tridiagonalisation in the step (eigh itself, and the eigh inside svd)
        ntrd = sum(1 for o in ops if o in ("eigh", "svd"))
        # (symmetric kernel: the lower triangle incl. diagonal, mt(mt+1)/2 elements)
        alg_bytes = ntrd * 8.0 * sum(float(n - c - 1) * (n - c) / 2 for c in range(n - 1))
        ach = alg_bytes / (kms.value * 1e-3) / 1e9
        cands.append({"kernel": "trd_symv_kernel<double>", "bound": "hbm", "achieved": ach, "peak": hbm_peak,
                      "unit": "GB/s", "frac": ach / hbm_peak, "traffic": None, "launches_per_step": kl.value,
                      "avg_launch_ms": kms.value / kl.value, "alg_bytes_per_launch": alg_bytes / kl.value,
                      "kernel_share_of_step": kms.value / step_ms, "peak_source": how})
    if gl.value > 0:
        pk = fp64_peak()
        ach = gflops / (gms.value * 1e-3) / 1e12
        cands.append({"kernel": "gemm_kernel<double> (DMMA m8n8k4)", "bound": "tensor", "achieved": ach,
                      "peak": pk[0], "unit": "TFLOP/s", "frac": ach / pk[0], "traffic": None,
                      "launches_per_step": gl.value, "avg_launch_ms": gms.value / gl.value,
                      "alg_flops_per_launch": gflops / gl.value, "kernel_share_of_step": gms.value / step_ms,
                      "peak_source": pk[1]})
    # DRAM traffic per launch of the same kernels from the committed ncu capture of this command
    # (tools/traffic_summ.py -> profiles/traffic.json); None when no capture exists for the kernel
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tk = json.load(f)["kernels"]
        for c in cands:
            key = "gemm_kernel" if c["kernel"].startswith("gemm_kernel") else "trd_symv_kernel"
            if key in tk and n == 8192 and set(ops) == {"eigh", "svd"}:
                c["traffic"] = tk[key]["dram_bytes_per_launch"]
                c["traffic_source"] = ("profiles/traffic.json: ncu dram__bytes_read+write, mean per launch over "
                                       + tk[key].get("sample", "one step"))
                if "alg_bytes_per_launch_same_sample" in tk[key]:
                    c["traffic_alg_bytes_same_launches"] = tk[key]["alg_bytes_per_launch_same_sample"]
    except (OSError, KeyError, ValueError):
        pass
    cands.sort(key=lambda c: -c["kernel_share_of_step"])
    roofline = cands[0] if cands else None
    if roofline and len(cands) > 1:
        roofline["second_kernel"] = cands[1]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # CPU baseline beside it (rank 0, N=1 only): bounded sample of the same workload
    cpu = None
    if world == 1 and not args.no_cpu:
        g, s, desc = cpu_sample(ops, args.cpu_n)
        cpu = {"value": g, "unit": "GFLOP/s", "cores": threads_info(), "kind": "port", "sample": desc, "seconds": s}
    line = {
        "metric": METRIC, "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
        "warmup": nwarm, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{'+'.join(o + ('_full!' if o == 'eigh' else '_compact!') for o in ops)} {n}x{n} Float64 "
                               "(BASELINE configs[1]); per-rank replica when n_gpus>1",
                   "l2": "inputs (512 MiB) larger than L2; A refreshed from a resident copy each step",
                   "factorizations_per_s": world * len(ops) * args.steps / (ms * 1e-3)},
        "e2e": {"value": e2e_value, "unit": "GFLOP/s", "h2d_bytes_per_step": bytes_io[0],
                "d2h_bytes_per_step": bytes_io[1], "ms_per_step": ms_e2e / e2e_steps},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--n", type=int, default=8192)
    ap.add_argument("--ops", default="eigh,svd")
    ap.add_argument("--cpu-n", type=int, default=4096, help="size of the bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
