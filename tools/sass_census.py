#!/usr/bin/env python
"""Opcode census of the product library: per kernel, how many tcgen05 MMA (UTCHMMA / UTCQMMA ...),
tensor-memory load/store (LDTM / STTM), TMA bulk copy / reduce (UBLKCP / UBLKRED / UTMA*), vector
reduction (RED) and MUFU instructions the SASS holds.  Evidence that the velocity MLP is tcgen05/TMEM
native and which kernels stream through TMA.

    python tools/sass_census.py [nvfi_b200/libnvfi_b200.so] > profiles/rNN_sass_census.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "nvfi_b200", "libnvfi_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
GROUPS = [("UTC*MMA", r"\bUTC[A-Z]*MMA"), ("UTCCP", r"\bUTCCP"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
          ("UBLKCP", r"\bUBLKCP"), ("UBLKRED", r"\bUBLKRED"), ("UTMA", r"\bUTMA"), ("SYNCS", r"\bSYNCS"),
          ("RED", r"\bRED\b|\bREDG"), ("ATOM", r"\bATOM"), ("MUFU", r"\bMUFU"), ("HMMA/IMMA (legacy)", r"\b[HI]MMA"),
          ("CCTL (discard)", r"\bCCTL"), ("STL/LDL (spill)", r"\b(STL|LDL)\b")]
kern = None
counts = collections.OrderedDict()
total = collections.Counter()
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        continue
    if kern is None or "/*" not in ln:
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if not m:
        continue
    op = m.group(1)
    counts[kern]["_n"] += 1
    for g, pat in GROUPS:
        if re.search(pat, op):
            counts[kern][g] += 1


def demangle(n):
    p = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    return re.sub(r"\(.*", "", p)


print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)} (sm_100a): instruction counts per kernel")
hdr = ["kernel", "instr"] + [g for g, _ in GROUPS]
rows = []
for k, c in counts.items():
    rows.append([demangle(k), str(c["_n"])] + [str(c[g]) if c[g] else "." for g, _ in GROUPS])
    for g, _ in GROUPS:
        total[g] += c[g]
rows.sort(key=lambda r: -int(r[1]))
w = [max(len(r[i]) for r in rows + [hdr]) for i in range(len(hdr))]
for r in [hdr] + rows:
    print("  ".join(x.ljust(w[i]) if i == 0 else x.rjust(w[i]) for i, x in enumerate(r)))
print("\n# totals: " + ", ".join(f"{g} {total[g]}" for g, _ in GROUPS))
