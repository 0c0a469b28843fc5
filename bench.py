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
the timed region).  After the timed region the LAST step's outputs are checked on the device (residual,
orthogonality, trace identities) and the figures go into the JSON line (`parity`).

N>1 (`--workload auto`): configs[1] does not shard (SURVEY 8e "replicas only"), so the multi-GPU line measures the
config that does: BASELINE configs[3], TSQR `qr_compact!` of a 16 777 216 x 256 Float64 matrix row-sharded over
the N ranks (STRONG scaling; one C-ABI call `makb200_tsqr(h, ncclComm_t, ...)` per step: local CholeskyQR2, binary
tree over NCCL on the R factors, tree factors folded into the last solve).  `--workload tsqr --gpus 1` runs the
same matrix on one GPU (the strong-scaling base); `--workload c2` forces per-rank replicas of configs[1].
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
def blas_all_cores():
    """torchrun exports OMP_NUM_THREADS=1: give the BLAS behind scipy every host core explicitly."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:
        pass


def cpu_sample(ops, n_cpu, reps=1):
    """time the oracle on a bounded sample; returns (GFLOP/s, seconds, description)."""
    from oracle import mak_oracle as O
    blas_all_cores()
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


def tsqr_flops(m, n):
    return 4.0 * m * n * n - 4.0 * n ** 3 / 3.0


def cpu_sample_tsqr(m_cpu, n, m_full):
    """oracle qr_compact (geqrt(36) + gemqrt on I) of an m_cpu x n row sample; GFLOP/s is per-row work, so the
    figure transfers linearly in m (the full 16.7M x 256 needs 69 GB and > 2^31 elements under LP64)."""
    from oracle import mak_oracle as O
    blas_all_cores()
    A = O.randn_matrix(m_cpu, n, "f64", seed=5)
    t0 = time.perf_counter()
    O.qr_compact(A)
    t = time.perf_counter() - t0
    return tsqr_flops(m_cpu, n) / t / 1e9, t, (f"qr_compact {m_cpu}x{n} f64 x1 (LAPACK geqrt+gemqrt via scipy/OpenBLAS); "
                                               f"extrapolated linearly in m to {m_full} rows")


def threads_info():
    try:
        from threadpoolctl import threadpool_info
        for lib in threadpool_info():
            if lib.get("user_api") == "blas":
                return int(lib.get("num_threads", os.cpu_count() or 1))
    except Exception:
        pass
    return os.cpu_count() or 1


def pick_workload(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.workload != "auto":
        return args.workload
    return "c2" if max(world, args.gpus) == 1 else "tsqr"


def run_reference(args):
    """The reference's CPU path (LAPACK replay = oracle, kind "port": Julia cannot run here) on ALL host cores.
    c2: ONE repetition at the bench's own n (8192: about 1.5 min) whatever --steps says - a smaller n would not be
    the same config; tsqr: a 2 097 152-row sample, linear in m."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = pick_workload(args)
    blas_all_cores()
    if wl == "tsqr":
        m_cpu = min(args.tsqr_rows, 1 << 21)
        cpu_sample_tsqr(1 << 16, args.tsqr_cols, args.tsqr_rows)
        v, secs, desc = cpu_sample_tsqr(m_cpu, args.tsqr_cols, args.tsqr_rows)
        ms_step = secs * 1e3 * args.tsqr_rows / m_cpu
        workload = (f"tsqr qr_compact! {args.tsqr_rows}x{args.tsqr_cols} Float64 (BASELINE configs[3]); reference step = "
                    f"{m_cpu}-row sample, ms_per_step extrapolated linearly in m")
        steps_done, scaling = 1, "strong"
    else:
        ops = args.ops.split(",")
        n_cpu = args.cpu_n if args.cpu_n > 0 else args.n
        cpu_sample(ops, min(n_cpu, 1024))
        v, secs, desc = cpu_sample(ops, n_cpu)
        ms_step = secs * 1e3
        extra = "" if n_cpu == args.n else f"; bounded sample n={n_cpu}, GFLOP/s extrapolated (flops ~ n^3)"
        workload = (f"{'+'.join(o + ('_full!' if o == 'eigh' else '_compact!') for o in ops)} {args.n}x{args.n} Float64 "
                    f"(BASELINE configs[1]); ONE repetition at n={n_cpu} independent of --steps{extra}")
        steps_done, scaling = 1, "weak"
    cores = threads_info()
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "GFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(ms_step),
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "repetitions_timed": steps_done, "host_cores": os.cpu_count()},
        "cpu_baseline": {"value": v, "unit": "GFLOP/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": v, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
def run_ours(args):
    if pick_workload(args) == "tsqr":
        return run_tsqr(args)
    return run_c2(args)


def run_c2(args):
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
    e2e_steps = max(1, args.steps)
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
        cands.append({"kernel": "trd_symv2_kernel<double> (persistent TMA-fed tridiagonalisation column kernel)", "bound": "hbm", "achieved": ach, "peak": hbm_peak,
                      "unit": "GB/s", "frac": ach / hbm_peak, "traffic": None, "launches_per_step": kl.value,
                      "avg_launch_ms": kms.value / kl.value, "alg_bytes_per_launch": alg_bytes / kl.value,
                      "kernel_share_of_step": kms.value / step_ms, "peak_source": how})
    if gl.value > 0:
        pk = fp64_peak()
        ach = gflops / (gms.value * 1e-3) / 1e12
        cands.append({"kernel": "gemm_tma_kernel<double> (TMA-fed DMMA m8n8k4; rank-k updates and unaligned operands: gemm_kernel)", "bound": "tensor", "achieved": ach,
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
            key = "gemm_kernel" if c["kernel"].startswith("gemm_") else "trd_symv_kernel"
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

    # ---- output check of the LAST step, on the device (torch ops are the checker here, not the product) ----
    parity = {"tolerance_10_n_eps": 10 * n * 2.220446049250313e-16}
    with torch.no_grad():
        eye = torch.eye(n, dtype=torch.float64, device=dev)
        if "eigh" in ops:
            D, V = outs["eigh"]
            w = D if D.dim() == 1 else torch.diagonal(D)
            H = dev_in["eigh"]
            nh = float(torch.linalg.matrix_norm(H))
            parity["eigh"] = {
                "resid_AV_VD_over_A": float(torch.linalg.matrix_norm(H @ V - V * w)) / nh,
                "orth_VhV_I": float(torch.linalg.matrix_norm(V.t() @ V - eye)),
                "trace_err": abs(float(w.sum() - torch.diagonal(H).sum())) / nh,
                "sum_lambda2_vs_fro2": abs(float((w * w).sum()) - nh * nh) / (nh * nh),
                "ascending": bool((w[1:] >= w[:-1]).all()),
            }
        if "svd" in ops:
            U, S, Vh = outs["svd"]
            G = dev_in["svd"]
            ng = float(torch.linalg.matrix_norm(G))
            res = float(torch.linalg.matrix_norm(G - (U * S) @ Vh)) / ng
            ou = float(torch.linalg.matrix_norm(U.t() @ U - eye))
            ov = float(torch.linalg.matrix_norm(Vh @ Vh.t() - eye))
            parity["svd"] = {
                "resid_A_USVh_over_A": res, "orth_UhU_I": ou, "orth_VhVhh_I": ov,
                "sum_sigma2_vs_fro2": abs(float((S * S).sum()) - ng * ng) / (ng * ng),
                "descending_nonneg": bool((S[1:] <= S[:-1]).all() and (S >= 0).all()),
                # Weyl: a factorization with this residual and orthogonality has |sigma - sigma_exact| <= bound * sigma_1
                "sigma_err_bound_over_sigma1": res * ng / float(S[0]) + ou + ov,
            }
        del eye
    parity["ok"] = all(v2 <= parity["tolerance_10_n_eps"] for k in ("eigh", "svd") if k in parity
                       for k2, v2 in parity[k].items() if k2.startswith(("resid", "orth")))
    parity["oracle_comparison"] = "tests/test_gpu_x_config_size.py compares sigma/lambda with the LAPACK oracle at n = 2048/4096"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # CPU baseline beside it (rank 0, N=1 only): bounded sample of the same workload
    cpu = None
    if world == 1 and not args.no_cpu:
        n_cpu = args.cpu_n if args.cpu_n > 0 else min(n, 4096)
        g, s, desc = cpu_sample(ops, n_cpu)
        if n_cpu != n:
            desc += f"; GFLOP/s extrapolated to n={n} (flops ~ n^3)"
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
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_tsqr(args):
    """BASELINE configs[3]: qr_compact! of a 16 777 216 x 256 Float64 matrix row-sharded over the ranks (strong scaling)."""
    import ctypes
    import torch
    import torch.distributed as dist
    import makb200
    from makb200 import tsqr as T

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = makb200._lib.load()
    M, n = args.tsqr_rows, args.tsqr_cols
    rows = [M // world + (1 if r < M % world else 0) for r in range(world)]
    m_loc = rows[rank]
    # synthetic shard, generated on the device with Philox, seed 5 + rank
    g = torch.Generator(device=dev)
    g.manual_seed(5 + rank)
    A_src = torch.randn((n, m_loc), dtype=torch.float64, device=dev, generator=g).t()
    A = makb200.colmajor_empty(m_loc, n, torch.float64, dev)
    Q = makb200.colmajor_empty(m_loc, n, torch.float64, dev)
    R = makb200.colmajor_empty(n, n, torch.float64, dev)
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    h = makb200.Handle.get(dev)
    comm = T.nccl_comm(None, dev) if world > 1 else ctypes.c_void_p(0)
    lw = lib.makb200_tsqr_worksize(h.h, 0, m_loc, n, world)
    work = torch.empty(max(int(lw), 1), dtype=torch.uint8, device=dev)

    def call():
        h2 = makb200.Handle.get(dev)
        rc = lib.makb200_tsqr(h2.h, comm, 0, m_loc, n, A.data_ptr(), m_loc, Q.data_ptr(), m_loc, R.data_ptr(), n,
                              work.data_ptr(), work.numel(), info.data_ptr())
        h2.check(rc, "makb200_tsqr")

    def step_dev():
        A.copy_(A_src)   # the factorization overwrites A: refresh from the resident copy (counted)
        call()

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

    nwarm = max(args.warmup, 3)
    for _ in range(nwarm):
        step_dev()
    torch.cuda.synchronize()
    if int(info.item()) != 0:
        raise RuntimeError("makb200_tsqr reported a Cholesky breakdown on the synthetic shard")
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = lib.makb200_launch_count()
    ms = timed(step_dev, args.steps)
    launches = int(lib.makb200_launch_count() - l0)
    clocks = sampler.stop() if rank == 0 else None
    fl = tsqr_flops(M, n)
    value = fl * args.steps / (ms * 1e-3) / 1e9
    ms_copy = timed(lambda: A.copy_(A_src), args.steps) / args.steps

    # ---- roofline of the dominant kernel: the DMMA GEMM (Gram products and triangular solves), one instrumented step
    lib.makb200_kernel_timing(1)
    step_dev()
    torch.cuda.synchronize()
    gms, gl = ctypes.c_double(), ctypes.c_int()
    lib.makb200_kernel_time(1, ctypes.byref(gms), ctypes.byref(gl))
    gflops = float(lib.makb200_gemm_flops())
    lib.makb200_kernel_timing(0)
    step_ms = ms / args.steps
    pk = fp64_peak()
    ach = gflops / (gms.value * 1e-3) / 1e12 if gms.value > 0 else 0.0
    roofline = {"kernel": "gemm_tma_kernel<double> (TMA-fed DMMA m8n8k4)", "bound": "tensor", "achieved": ach, "peak": pk[0],
                "unit": "TFLOP/s", "frac": ach / pk[0], "traffic": None, "launches_per_step": gl.value,
                "avg_launch_ms": gms.value / max(gl.value, 1), "alg_flops_per_launch": gflops / max(gl.value, 1),
                "kernel_share_of_step": gms.value / step_ms, "peak_source": pk[1],
                "note": "flops = executed DMMA tiles of this rank's launches (lower-mode launches count executed tiles only)"}

    # ---- parity of the last step on the device: residual on a row sample, global orthogonality, R^H R = A^H A ----
    with torch.no_grad():
        idx = torch.randint(0, m_loc, (min(m_loc, 65536),), device=dev)
        num = torch.linalg.matrix_norm(A_src[idx] - Q[idx] @ R) ** 2
        den = torch.linalg.matrix_norm(A_src[idx]) ** 2
        Gq = torch.zeros((n, n), dtype=torch.float64, device=dev)
        Ga = torch.zeros((n, n), dtype=torch.float64, device=dev)
        for r0 in range(0, m_loc, 1 << 20):
            qs, as_ = Q[r0:r0 + (1 << 20)], A_src[r0:r0 + (1 << 20)]
            Gq += qs.t() @ qs
            Ga += as_.t() @ as_
        red = torch.stack([num, den])
        if world > 1:
            dist.all_reduce(red)
            dist.all_reduce(Gq)
            dist.all_reduce(Ga)
        tolp = 10 * M * 2.220446049250313e-16
        parity = {
            "tolerance_10_n_eps": tolp,
            "resid_A_QR_over_A_row_sample": float(torch.sqrt(red[0] / red[1])),
            "orth_QhQ_I_global": float(torch.linalg.matrix_norm(Gq - torch.eye(n, dtype=torch.float64, device=dev))),
            "gram_identity_RhR_vs_AhA": float(torch.linalg.matrix_norm(R.t() @ R - Ga) / torch.linalg.matrix_norm(Ga)),
            "diagR_positive_upper": bool((torch.diagonal(R) > 0).all() and float(torch.tril(R, -1).abs().max()) == 0.0),
            "oracle_comparison": "tests/test_gpu_tsqr_multi.py (torchrun) compares Q_p and R with the LAPACK oracle of the concatenated matrix",
        }
        parity["ok"] = (parity["resid_A_QR_over_A_row_sample"] <= tolp and parity["orth_QhQ_I_global"] <= tolp
                        and parity["diagR_positive_upper"])
        del Gq, Ga

    # ---- e2e: HOST shards in, HOST Q and R out, through the same C-ABI call.  Pinned staging of `chunk` rows is
    # reused for every slab of the shard (bounded host memory), so the device sees a shard of repeated row slabs.
    chunk = min(m_loc, 1 << 19)
    hin = torch.empty((n, chunk), dtype=torch.float64).pin_memory()
    hin.copy_(A_src[:chunk].t())
    hq = torch.empty((n, chunk), dtype=torch.float64).pin_memory()
    hr = torch.empty((n, n), dtype=torch.float64).pin_memory()
    bytes_io = [0, 0]

    def step_e2e():
        h2d = d2h = 0
        for r0 in range(0, m_loc, chunk):
            c = min(chunk, m_loc - r0)
            A[r0:r0 + c].t().copy_(hin[:, :c], non_blocking=True)
            h2d += c * n * 8
        call()
        for r0 in range(0, m_loc, chunk):
            c = min(chunk, m_loc - r0)
            hq[:, :c].copy_(Q[r0:r0 + c].t(), non_blocking=True)
            d2h += c * n * 8
        hr.copy_(R.t(), non_blocking=True)
        d2h += n * n * 8
        bytes_io[0], bytes_io[1] = h2d, d2h
    step_e2e()
    e2e_steps = max(1, min(args.steps, 3))
    ms_e2e = timed(step_e2e, e2e_steps)
    e2e_value = fl * e2e_steps / (ms_e2e * 1e-3) / 1e9

    if rank != 0:
        if world > 1:
            dist.barrier()        # rank 0 measures the one-GPU time of the same matrix before everybody leaves
            dist.destroy_process_group()
        return
    # ---- the SAME workload on ONE GPU (rank 0, after the timed regions): the base of the strong-scaling curve.
    # `bench.py --gpus 1` runs configs[1] (the N = 1 headline), so the N > 1 line carries its own one-GPU reference.
    one_gpu = None
    if world > 1 and not args.no_one_gpu_ref:
        try:
            del A_src, A, Q, work, hin, hq
            torch.cuda.empty_cache()
            g1 = torch.Generator(device=dev)
            g1.manual_seed(5)
            S1 = torch.randn((n, M), dtype=torch.float64, device=dev, generator=g1).t()
            A1 = makb200.colmajor_empty(M, n, torch.float64, dev)
            Q1 = makb200.colmajor_empty(M, n, torch.float64, dev)
            lw1 = lib.makb200_tsqr_worksize(h.h, 0, M, n, 1)
            w1 = torch.empty(max(int(lw1), 1), dtype=torch.uint8, device=dev)

            def step1():
                A1.copy_(S1)
                h2 = makb200.Handle.get(dev)
                rc = lib.makb200_tsqr(h2.h, ctypes.c_void_p(0), 0, M, n, A1.data_ptr(), M, Q1.data_ptr(), M, R.data_ptr(), n,
                                      w1.data_ptr(), w1.numel(), info.data_ptr())
                h2.check(rc, "makb200_tsqr")
            for _ in range(2):
                step1()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                step1()
            e1.record()
            torch.cuda.synchronize()
            ms1 = e0.elapsed_time(e1) / args.steps
            one_gpu = {"ms_per_step": ms1, "value": fl / (ms1 * 1e-3) / 1e9, "unit": "GFLOP/s",
                       "note": "same 16.7M x 256 matrix, one rank, measured by rank 0 in this run after the timed regions"}
            del S1, A1, Q1, w1
        except Exception as exc:  # noqa: BLE001  (out of memory on a smaller device: report, do not fail the line)
            one_gpu = {"unavailable": repr(exc)[:200]}
    if world > 1:
        dist.barrier()
    cpu = None
    if world == 1 and not args.no_cpu:
        gcpu, scpu, desc = cpu_sample_tsqr(min(M, 1 << 20), n, M)
        cpu = {"value": gcpu, "unit": "GFLOP/s", "cores": threads_info(), "kind": "port", "sample": desc, "seconds": scpu}
    api_from = None
    line = {
        "metric": METRIC, "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
        "warmup": nwarm, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"tsqr qr_compact! {M}x{n} Float64 row-sharded over {world} rank(s) (BASELINE configs[3])",
                   "rows_per_rank": rows[0], "flops_per_step": fl,
                   "l2": f"inputs ({m_loc * n * 8 / 2**30:.1f} GiB per rank) larger than L2; A refreshed from a resident copy each step "
                         f"({ms_copy:.2f} ms of the step)",
                   "collective": ("binary tree over ranks: ncclSend/ncclRecv of one n x n R factor per round on the compute stream, "
                                  "ncclBroadcast of R; own communicator from makb200_comm_create") if world > 1 else "none (1 rank)",
                   "factorizations_per_s": args.steps / (ms * 1e-3),
                   "hbm_floor_ms_per_rank": 16.0 * m_loc * n / (peaks()[0] * 1e9) * 1e3,
                   "same_workload_on_one_gpu": one_gpu},
        "e2e": {"value": e2e_value, "unit": "GFLOP/s", "h2d_bytes_per_step": bytes_io[0] * world,
                "d2h_bytes_per_step": bytes_io[1] * world, "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps,
                "host_buffer": f"pinned staging of {chunk} rows reused per slab"},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
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
    ap.add_argument("--cpu-n", type=int, default=0, help="n of the CPU sample (0: reference arm = --n, cpu_baseline = min(n, 4096))")
    ap.add_argument("--workload", default="auto", choices=["auto", "c2", "tsqr"],
                    help="auto: configs[1] (eigh+svd 8192^2) at 1 GPU, configs[3] (TSQR 16.7M x 256, strong scaling) at N > 1")
    ap.add_argument("--tsqr-rows", type=int, default=16777216)
    ap.add_argument("--tsqr-cols", type=int, default=256)
    ap.add_argument("--no-one-gpu-ref", action="store_true", help="N > 1: skip the one-GPU run of the same TSQR matrix on rank 0")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
