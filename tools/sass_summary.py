#!/usr/bin/env python
"""Per-kernel SASS opcode summary of the built library: which kernels carry tcgen05 MMAs
(`UTCHMMA` / `UTCQMMA`), TMEM loads (`LDTM`), bulk copies (`UBLKCP`), packed FP32 (`FFMA2`), legacy
tensor-core `HMMA`, plus registers from the ELF.  Usage: python tools/sass_summary.py [lib.so] > profiles/sass_summary.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "retisar_b200/lib/libretisar_b200.so"
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "UBLKCP", "SYNCS", "FFMA2", "FFMA", "FADD2", "FMUL2", "HMMA",
         "LDGSTS", "LDG", "STG", "LDS", "STS", "SHFL", "MUFU", "BAR"]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
per = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = cur.replace("(anonymous namespace)::", "").replace("rtb::", "")
        cur = re.sub(r"\(.*", "", cur)
        per[cur] = collections.Counter()
        continue
    if cur:
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_]+)", line)
        if m:
            per[cur]["total"] += 1
            op = m.group(2)
            for w in WATCH:
                if op == w or op.startswith(w + "_") or (w in ("UTCHMMA", "UTCQMMA", "LDTM", "UBLKCP", "LDGSTS", "SYNCS") and op.startswith(w)):
                    per[cur][w] += 1
                    break
tot = collections.Counter()
print(f"# SASS opcode counts per kernel of {lib} (static; cuobjdump -sass)")
print(f"{'kernel':70s} {'total':>7s} " + " ".join(f"{w:>7s}" for w in WATCH))
for k, c in sorted(per.items()):
    if not c["total"]:
        continue
    print(f"{k[:70]:70s} {c['total']:7d} " + " ".join(f"{c[w]:7d}" for w in WATCH))
    tot.update(c)
print(f"{'ALL':70s} {tot['total']:7d} " + " ".join(f"{tot[w]:7d}" for w in WATCH))
