#!/usr/bin/env python
"""Per-kernel SASS mnemonic census of libspann3r_b200.so (`cuobjdump -sass`): the tcgen05 / TMEM / TMA instructions that
prove a Blackwell-native kernel (B200_PROFILING.md: UTC*MMA, LDTM / STTM, UTMALDG / UTMASTG / UBLKCP) and the legacy ones
that must be absent (HMMA, HGMMA).  Runs without a GPU.  Usage: python tools/sass_evidence.py > profiles/rN_sass_mnemonics.md"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "spann3r_b200", "libspann3r_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
WATCH = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "ELECT", "MUFU",
         "HMMA", "HGMMA", "QGMMA", "IGMMA", "LDGSTS"]
kern = OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kern[cur] = Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m:
        op = m.group(1)
        kern[cur]["_total"] += 1
        for w in WATCH:
            if op.startswith(w):
                kern[cur][w] += 1
                break
names = subprocess.run(["c++filt"] + list(kern), capture_output=True, text=True).stdout.splitlines()
print("# SASS mnemonic census of libspann3r_b200.so (cuobjdump -sass, sm_100a)\n")
print("Counts of static instructions per kernel.  `UTCHMMA` = tcgen05.mma (kind::f16 / tf32), `LDTM` / `STTM` = tcgen05.ld / st,\n"
      "`UTMALDG` = TMA tensor load, `UTCBAR` = tcgen05.commit, `SYNCS` = mbarrier ops.  `HMMA` / `HGMMA` (legacy mma.sync / wgmma)\n"
      "must be 0 everywhere.\n")
cols = [w for w in WATCH if any(k[w] for k in kern.values())] + ["HMMA", "HGMMA"]
cols = list(OrderedDict.fromkeys(cols))
print("| kernel | instr | " + " | ".join(cols) + " |")
print("|---|---:|" + "---:|" * len(cols))
for (mangled, c), name in zip(kern.items(), names):
    short = re.sub(r"\(.*", "", name).replace("void ", "")
    print(f"| `{short[:70]}` | {c['_total']} | " + " | ".join(str(c[w]) for w in cols) + " |")
tot = Counter()
for c in kern.values():
    tot.update(c)
print(f"\n{len(kern)} kernels; totals: " + ", ".join(f"{w} {tot[w]}" for w in cols))
assert tot["HMMA"] == 0 and tot["HGMMA"] == 0, "legacy tensor-core instructions found"
