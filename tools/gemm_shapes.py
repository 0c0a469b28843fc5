"""Aggregate a MAKB200_GEMM_LOG file (one line per GEMM launch of an instrumented step: m n k flags flops ms)
by shape class: where the DMMA GEMM time of a step goes and at what rate."""
import sys
from collections import defaultdict

rows = [l.split() for l in open(sys.argv[1]) if l.strip()]
agg = defaultdict(lambda: [0, 0.0, 0.0])
for m, n, k, fl, flops, ms in rows:
    m, n, k, fl = int(m), int(n), int(k), int(fl)
    def cls(x):
        return x if x <= 256 else (512 if x <= 512 else (1024 if x <= 1024 else (2048 if x <= 2048 else (4096 if x <= 4096 else 8192))))
    key = (("T" if fl & 1 else "N") + ("T" if fl & 2 else "N") + (" lower" if fl & 4 else "") + (" 2cta" if fl & 8 else "") + (f" splitk{fl >> 4}" if (fl >> 4) > 1 else ""),
           cls(m), cls(n), cls(k))
    a = agg[key]
    a[0] += 1; a[1] += float(flops); a[2] += float(ms)
tot = sum(a[2] for a in agg.values())
totf = sum(a[1] for a in agg.values())
print(f"{len(rows)} launches, {tot:.1f} ms, {totf / tot / 1e9:.2f} TF/s overall")
print(f"{'ops':18s} {'m<=':>6s} {'n<=':>6s} {'k<=':>6s} {'count':>6s} {'ms':>9s} {'share':>6s} {'TF/s':>7s}")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:40]:
    print(f"{key[0]:18s} {key[1]:6d} {key[2]:6d} {key[3]:6d} {a[0]:6d} {a[2]:9.2f} {100 * a[2] / tot:5.1f}% {a[1] / a[2] / 1e9:7.2f}")
