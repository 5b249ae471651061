"""Turn an .ncu-rep into the text summaries kept under profiles/ (details metrics + stall / opcode / per-barrier-segment
breakdown of the hot loop from the source page)."""
import csv, io, re, subprocess, sys
from collections import Counter
rep, out = sys.argv[1], sys.argv[2]
kid = sys.argv[3] if len(sys.argv) > 3 else None
def page(p):
    return subprocess.run(["ncu", "-i", rep, "--page", p, "--csv"], capture_output=True, text=True).stdout
det = list(csv.reader(io.StringIO(page("details"))))
hi = next(i for i, r in enumerate(det) if "Metric Name" in r)
h = det[hi]; ix = {n: i for i, n in enumerate(h)}
lines = []
seen = set()
for r in det[hi + 1:]:
    if len(r) < ix["Metric Value"] + 1: continue
    if kid is not None and r[ix["ID"]] != kid: continue
    sec, name, val, unit = r[ix["Section Name"]], r[ix["Metric Name"]], r[ix["Metric Value"]], r[ix["Metric Unit"]]
    if not name or (r[ix["ID"]], sec, name) in seen: continue
    seen.add((r[ix["ID"]], sec, name))
    lines.append(f"[{r[ix['ID']]}] {r[ix['Kernel Name']][:48]:48s} | {sec} | {name} | {val} {unit}")
open(out, "w").write("\n".join(lines) + "\n")
print(out, len(lines), "metric lines")
