"""CPU baseline for BASELINE configs[2] (batched block-sparse qr_compact! / svd_trunc! of 20 000 ComplexF64 blocks,
sizes 16..512): the reference has no batched entry point, so the baseline is its per-block call in a loop
on all host cores - here the oracle's LAPACK replay (geqrt+gemqrt / gesdd + slice), one LAPACK call per worker thread
at a time (BLAS threads = 1 inside a worker, `threads` workers; scipy's LAPACK releases the GIL).
A bounded sample per size bucket is timed and scaled to the bucket's block count; the JSON says so.
  python tools/c3_cpu_baseline.py [sample_per_bucket] [out.json]"""
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import mak_oracle as O


def main():
    sample = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/r2_c3_cpu_baseline.json"
    threads = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=1)
    except Exception:
        pass
    rng = np.random.Generator(np.random.PCG64(4))
    dims = np.rint(16 * 32 ** rng.random(20000)).astype(int)
    buckets = [(16, 32), (33, 64), (65, 128), (129, 256), (257, 512)]
    res = {"workload": "20000 ComplexF64 blocks, n = round(16*32^u) (SURVEY 8d), per-block LAPACK loop", "threads": threads,
           "blas_threads_per_worker": 1, "sample_per_bucket": sample, "buckets": {}}
    tot = {"qr": 0.0, "svd_trunc": 0.0}
    for lo, hi in buckets:
        idx = np.nonzero((dims >= lo) & (dims <= hi))[0]
        pick = idx[:: max(1, len(idx) // sample)][:sample]
        blocks = [O.randn_matrix(int(dims[i]), int(dims[i]), "c128", seed=4000 + int(i)) for i in pick]
        for op, fn in (("qr", lambda a: O.qr_compact(a)), ("svd_trunc", lambda a: O.svd_trunc(a, O.truncrank(a.shape[0] // 2)))):
            with ThreadPoolExecutor(max_workers=threads) as ex:
                list(ex.map(fn, blocks[: min(len(blocks), threads)]))     # warm-up
                t0 = time.perf_counter()
                list(ex.map(fn, blocks))
                dt = time.perf_counter() - t0
            rate = len(blocks) / dt
            full = len(idx) / rate
            tot[op] += full
            res["buckets"][f"{op}_{lo}-{hi}"] = {"blocks_in_config": int(len(idx)), "sampled": len(blocks), "sample_seconds": dt,
                                                 "blocks_per_s": rate, "seconds_for_bucket_extrapolated": full}
            print(f"{op:10s} {lo:3d}-{hi:3d}: {rate:10.1f} blocks/s  -> {full:7.2f} s for {len(idx)} blocks", flush=True)
    res["total_seconds_extrapolated"] = tot
    os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
    with open(out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(tot))


if __name__ == "__main__":
    main()
