#!/usr/bin/env python3
"""SASS opcode summary of libpmm_b200.so per kernel: which tensor-core / TMA / TMEM instructions each one contains.
    python scripts/sass_summary.py > profiles/sass_r2.md          (cuobjdump only; no GPU needed)
UTCHMMA = tcgen05.mma (kind::f16 / tf32), UTMALDG / UTMASTG = TMA tensor load / store, LDTM = tcgen05.ld (TMEM -> registers),
UTCBAR = tcgen05.commit, DMMA = mma.sync f64, HMMA/IMMA would be legacy mma.sync (none expected)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "polars_matmul_b200", "libpmm_b200.so")
OPS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "UTCBAR", "UTCATOMSWS", "DMMA", "HMMA", "IMMA", "SYNCS", "REDUX", "SHFL", "FFMA", "DFMA"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
cur, counts = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1).split(".")[0]
        counts[cur]["_total"] += 1
        if op in OPS:
            counts[cur][op] += 1
            if op in ("UTCHMMA", "UTMALDG") and ".2CTA" in m.group(1):
                counts[cur][op + ".2CTA"] += 1
print("# SASS opcode summary per kernel (round 2)\n")
print("`cuobjdump -sass polars_matmul_b200/libpmm_b200.so` (sm_100a), counted by `scripts/sass_summary.py`. UTCHMMA = tcgen05.mma,")
print("UTMALDG/UTMASTG = TMA tensor load/store, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, DMMA = FP64 mma.sync. No HMMA/IMMA (legacy")
print("mma.sync) anywhere: the f32/f16 contraction is tcgen05 only.\n")
cols = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMALDG.2CTA", "UTMASTG", "LDTM", "UTCBAR", "DMMA", "HMMA", "IMMA", "FFMA", "DFMA", "SHFL"]
print("| kernel | instr | " + " | ".join(cols) + " |")
print("|---|---|" + "---|" * len(cols))
tot = collections.Counter()
for fn, c in counts.items():
    name = demangle(fn)
    name = re.sub(r"pmm::\(anonymous namespace\)::", "", name)
    name = re.sub(r"\(.*", "", name)[:70]
    print(f"| `{name}` | {c['_total']} | " + " | ".join(str(c.get(k, 0) or "") for k in cols) + " |")
    tot.update(c)
print(f"| **all {len(counts)} kernels** | {tot['_total']} | " + " | ".join(str(tot.get(k, 0) or "") for k in cols) + " |")
