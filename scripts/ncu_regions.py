"""Attribute executed warp instructions of an ncu report to source-line ranges.

    python scripts/ncu_regions.py rep.ncu-rep file.cuh name:lo-hi name:lo-hi ...
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep, fname = sys.argv[1], sys.argv[2]
regions = []
for a in sys.argv[3:]:
    n, r = a.split(":")
    lo, hi = r.split("-")
    regions.append((n, int(lo), int(hi)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
per = defaultdict(int)
ops = defaultdict(int)
cur, curf, hdr = None, "", None
for row in csv.reader(io.StringIO(raw)):
    if not row:
        continue
    if row[0] == "File Path":
        curf = row[1].split("/")[-1]
        continue
    if row[0] == "Line No":
        hdr = row
        ii = hdr.index("Instructions Executed")
        continue
    if hdr is None or len(row) <= ii:
        continue
    if row[0]:
        cur = (curf, int(row[0]))
    elif cur:
        try:
            n = int(row[ii])
        except ValueError:
            continue
        per[cur] += n
        tok = row[3].split()
        op = tok[1] if tok and tok[0].startswith("@") and len(tok) > 1 else (tok[0] if tok else "?")
        ops[op.split(".")[0]] += n
tot = sum(per.values()) or 1
print("total", tot)
for name, lo, hi in regions:
    n = sum(v for (f, l), v in per.items() if f == fname and lo <= l <= hi)
    print(f"  {name:18s} {100 * n / tot:5.1f}%  {n}")
other = sum(v for (f, l), v in per.items() if f != fname)
print(f"  other files        {100 * other / tot:5.1f}%")
print("  top opcodes:", ", ".join(f"{k}:{100 * v / tot:.1f}%" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:18]))
