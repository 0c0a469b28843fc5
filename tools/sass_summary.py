"""Static SASS evidence for every kernel of libmakb200.so (runs on the CPU box: `cuobjdump -sass`).
Per kernel: instruction count and the counts of the mnemonics that identify the hardware path —
DMMA (FP64 tensor core, mma.sync.m8n8k4.f64), LDGSTS (cp.async), UTMALDG/UBLKCP (TMA), DFMA/DADD/DMUL
(FP64 vector), LDS/STS, SHFL, BAR, cluster barriers, griddepcontrol (ACQBULK/PDL shows as .. ).
usage: python tools/sass_summary.py [lib.so] > profiles/r1_sass_static.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "matrixalgebrakit.jl_b200/libmakb200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = {}
fn = None
rows = collections.OrderedDict()
ins_re = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)")
for line in txt.splitlines():
    s = line.strip()
    if s.startswith("Function :"):
        fn = s.split(":", 1)[1].strip()
        rows[fn] = collections.Counter()
        continue
    m = ins_re.match(line)
    if m and fn:
        op = m.group(1)
        c = rows[fn]
        c["n"] += 1
        if op == "DMMA": c["DMMA"] += 1
        elif op == "LDGSTS": c["LDGSTS"] += 1
        elif op.startswith("UTMA") or op == "UBLKCP": c["TMA"] += 1
        elif op in ("DFMA", "DADD", "DMUL"): c["FP64"] += 1
        elif op in ("LDS", "STS", "LDSM"): c["SMEM"] += 1
        elif op == "SHFL": c["SHFL"] += 1
        elif op == "BAR": c["BAR"] += 1
        elif "CGA" in op or op in ("UCGABAR_ARV", "UCGABAR_WAIT"): c["CLUSTER"] += 1
        elif op in ("ATOM", "ATOMG", "RED", "ATOMS"): c["ATOM"] += 1
        elif op == "ACQBULK" or op.startswith("SYNCS"): c["MBAR"] += 1
names = list(rows)
try:
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    demangle = dict(zip(names, dem))
except Exception:
    pass
cols = ["n", "DMMA", "LDGSTS", "TMA", "FP64", "SMEM", "SHFL", "BAR", "CLUSTER", "ATOM", "MBAR"]
print("# static SASS summary of", lib, "(sm_100a; cuobjdump -sass); columns:", " ".join(cols), "kernel")
for fn, c in sorted(rows.items(), key=lambda kv: demangle.get(kv[0], kv[0])):
    d = demangle.get(fn, fn)
    d = re.sub(r"\(.*", "", d)
    d = d.replace("mak::", "").replace("void ", "")
    print(" ".join(f"{c[k]:6d}" for k in cols), d)
tot = collections.Counter()
for c in rows.values():
    tot.update(c)
print("# kernels:", len(rows), " total:", " ".join(f"{k}={tot[k]}" for k in cols))
