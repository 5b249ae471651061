"""Per-barrier-segment breakdown of an ncu source page csv: python src_segments.py src.csv [min_exec]"""
import csv, sys, re
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]; ix = {n: i for i, n in enumerate(h)}
min_exec = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
seg = []; cur = None
tot_samples = 0
for r in rows[hi + 1:]:
    if len(r) < len(h): continue
    ex = int(r[ix["Instructions Executed"]] or 0); smp = int(r[ix["# Samples"]] or 0)
    tot_samples += smp
    sass = r[ix["Source"]].strip()
    if cur is None or ex != cur["ex"] and abs(ex - cur["ex"]) > 0.02 * max(ex, cur["ex"]):
        cur = {"ex": ex, "n": 0, "smp": 0, "st": Counter(), "ops": Counter(), "first": sass, "addr": r[ix["Address"]]}
        seg.append(cur)
    cur["n"] += 1; cur["smp"] += smp
    op = sass.split()[0] if not sass.startswith("@") else sass.split()[1]
    cur["ops"][op.split(".")[0]] += 1
    for c in stall_cols:
        v = int(r[ix[c]] or 0)
        if v: cur["st"][c[6:]] += v
    if "BAR" in sass or "WARPSYNC" in sass:
        cur = None
print("total samples", tot_samples)
for s in seg:
    if s["ex"] < min_exec or s["smp"] == 0: continue
    st = ", ".join(f"{k}:{v}" for k, v in s["st"].most_common(5))
    ops = ", ".join(f"{k}:{v}" for k, v in s["ops"].most_common(6))
    print(f"{s['addr'][-6:]} ex={s['ex']:6d} n={s['n']:4d} samples={s['smp']:6d} ({100*s['smp']/tot_samples:4.1f}%) | {st} | {ops}")
