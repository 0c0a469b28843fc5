"""BASELINE configs[2]: batched block-sparse qr_compact! / svd_trunc! of ComplexF64 blocks, sizes 16-512
(log-uniform), distributed over ranks by LPT (longest processing time first).
  python tools/batched_bench.py [nblocks] [maxdim]           (1 GPU)
  torchrun ... tools/batched_bench.py [nblocks] [maxdim]     (N GPUs, no data-path collective)
Reports per size bucket: blocks/s, algorithmic GB/s (HBM fraction) and GFLOP/s."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import makb200

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
nblocks = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
maxdim = int(sys.argv[2]) if len(sys.argv) > 2 else 512
ops = (sys.argv[3] if len(sys.argv) > 3 else "qr,svd").split(",")

rng = np.random.Generator(np.random.PCG64(4))
dims = np.rint(16 * 32 ** rng.random(nblocks)).astype(int)
dims = dims[dims <= maxdim]
# LPT partition (makb200.partition): no data-path collective
owner, imbalance = makb200.lpt_partition([(int(n), int(n)) for n in dims], world)
mine = dims[owner == rank]
g = torch.Generator(device=dev); g.manual_seed(4 + rank)
As0 = [torch.randn((n, n), dtype=torch.complex128, device=dev, generator=g).t() for n in mine]
buckets = [(16, 32), (33, 64), (65, 128), (129, 256), (257, 512)]


def run(op, idx, rep=1):
    blocks = [As0[i] for i in idx] * rep
    if op == "eigh":
        blocks = [(a + a.conj().t()) for a in blocks]
    As = [makb200.colmajor_empty(a.shape[0], a.shape[1], a.dtype, dev) for a in blocks]
    if op == "qr":
        plan = makb200.BatchedQRPlan(As)     # argument arrays built once (sizes/pointers are fixed)
    elif op == "eigh":
        plan = makb200.BatchedEighPlan(As)
    else:
        plan = makb200.BatchedSVDPlan(As)
    fn = plan.run
    if op == "svdtrunc":
        # config 3 as stated: svd_trunc!(trunc = truncrank(n_i // 2)) per block = batched compact SVD + ONE
        # truncation launch + one read; the kept factors are views
        spec = makb200.truncation.device_spec(makb200.notrunc())
        caps = [a.shape[0] // 2 for a in As]

        def fn():
            outs = plan.run()
            makb200.truncation.trunc_select_batched_([S for _, S, _ in outs], spec, caps)
    ts = []
    for it in range(3):
        for a, b in zip(As, blocks):
            a.copy_(b)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts[1:]))


res = {}
for op in ops:
    for lo, hi in buckets:
        idx = [i for i, n in enumerate(mine) if lo <= n <= hi]
        if not idx:
            continue
        cap = int(os.environ.get("MAKB200_BENCH_BIG_CAP", "64"))
        if op in ("svd", "svdtrunc", "eigh") and lo > 64 and len(idx) > cap:
            idx = idx[:cap]         # large-block SVD/eigh go through the single-matrix path: bounded sample (env lifts it)
        # the smallest bucket is replicated so that the launch carries enough bytes to show the
        # steady-state HBM fraction (4151 blocks are only ~0.1 GB = 17 us at HBM speed)
        rep = 16 if (hi <= 32 and len(idx) * 16 <= 80000) else 1
        ms = run(op, idx, rep)
        ns = np.array([mine[i] for i in idx] * rep, dtype=np.float64)
        if op == "qr":
            byt = 16 * 3 * (ns ** 2).sum()
            fl = 4 * (8.0 / 3) * (ns ** 3).sum()
        elif op in ("svd", "svdtrunc"):
            byt = 16 * 3 * (ns ** 2).sum() + 8 * ns.sum()
            fl = 4 * (20.0 / 3) * (ns ** 3).sum()
        else:
            byt = 16 * 2 * (ns ** 2).sum() + 8 * ns.sum()
            fl = 4 * (10.0 / 3) * (ns ** 3).sum()
        res[f"{op}_{lo}-{hi}"] = {"blocks": len(ns), "ms": ms, "blocks_per_s": len(ns) / ms * 1e3,
                                  "alg_GBs": byt / ms / 1e6, "hbm_frac": byt / ms / 1e6 / 6555.8,
                                  "alg_GFLOPs": fl / ms / 1e6, "replicated": rep}
if world > 1:
    gathered = [None] * world
    dist.all_gather_object(gathered, res)
else:
    gathered = [res]
if rank == 0:
    agg = {}
    for k in gathered[0]:
        ms = max(g[k]["ms"] for g in gathered if k in g)
        nb = sum(g[k]["blocks"] for g in gathered if k in g)
        agg[k] = {"blocks": nb, "ms_max_over_ranks": ms, "blocks_per_s": nb / ms * 1e3,
                  "alg_GBs": sum(g[k]["alg_GBs"] * g[k]["ms"] for g in gathered if k in g) / ms,
                  "alg_GFLOPs": sum(g[k]["alg_GFLOPs"] * g[k]["ms"] for g in gathered if k in g) / ms}
        agg[k]["hbm_frac_per_gpu"] = agg[k]["alg_GBs"] / world / 6555.8
    print(json.dumps({"workload": f"batched c128 blocks n={len(dims)} dims 16-{maxdim}", "n_gpus": world,
                      "lpt_imbalance": imbalance, "buckets": agg}, indent=1))
if world > 1:
    dist.destroy_process_group()
