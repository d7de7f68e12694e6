#!/usr/bin/env python
"""Hot SASS instructions of one kernel from an ncu report captured with --set full --import-source on.

    python tools/ncu_hot.py gpurun_out/X.ncu-rep k_advect_bwd_tc [top_n]

Prints (a) stall samples aggregated per code segment (split at BAR / mbarrier waits), (b) the top
instructions by stall samples with their dominant stall reasons."""
import csv
import io
import subprocess
import sys

path, kern = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
lines = out.split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
# a regex that matches several launches repeats the header: keep the first launch only
for i, r in enumerate(rows):
    if r.get("Address") == "Address":
        rows = rows[:i]
        break
stall_cols = [c for c in rows[0].keys() if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(r["# Samples"] or 0) for r in rows)
print(f"# {kern}: {len(rows)} instructions, {tot} stall samples")
seg, acc, seg_start = [], 0, 0
segst = {c: 0 for c in stall_cols}
for i, r in enumerate(rows):
    n = int(r["# Samples"] or 0)
    acc += n
    for c in stall_cols:
        segst[c] += int(r[c] or 0)
    s = r["Source"].strip()
    if s.startswith(("BAR", "SYNCS.PHASECHK", "@P0 BAR", "WARPSYNC")) or "SYNCS.PHASECHK" in s or " BAR." in s:
        seg.append((seg_start, i, acc, dict(segst), s[:40]))
        acc, seg_start, segst = 0, i + 1, {c: 0 for c in stall_cols}
seg.append((seg_start, len(rows) - 1, acc, dict(segst), "end"))
print("# segments (instruction index range, samples, share, top stalls, closing instruction)")
for a, b, n, st, s in seg:
    if n * 200 > tot:
        top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        print(f"  [{a:5d},{b:5d}] {n:8d} {100*n/tot:5.1f}%  " + " ".join(f"{k[6:]}={v}" for k, v in top) + f"   | {s}")
print("# top instructions")
order = sorted(range(len(rows)), key=lambda i: -int(rows[i]["# Samples"] or 0))[:top_n]
for i in sorted(order):
    r = rows[i]
    n = int(r["# Samples"] or 0)
    top = sorted(((int(r[c] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    print(f"  {i:5d} {n:7d} {100*n/tot:5.1f}%  {r['Source'].strip()[:70]:70s} " + " ".join(f"{k}={v}" for v, k in top))
if len(sys.argv) > 5:
    a, b = int(sys.argv[4]), int(sys.argv[5])
    print(f"# range {a}..{b}")
    for i in range(a, b + 1):
        r = rows[i]
        n = int(r["# Samples"] or 0)
        top = sorted(((int(r[c] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
        print(f"  {i:5d} {n:7d}  {r['Source'].strip()[:80]:80s} " + " ".join(f"{k}={v}" for v, k in top if v))
