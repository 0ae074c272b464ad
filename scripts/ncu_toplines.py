"""Top source lines of an ncu report by executed warp instructions and by stall samples.

    python scripts/ncu_toplines.py rep.ncu-rep [n_lines] [kernel-id]
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if len(sys.argv) > 3:
    cmd += ["--kernel-id", sys.argv[3]]
raw = subprocess.run(cmd, capture_output=True, text=True).stdout
inst = defaultdict(int)
samp = defaultdict(int)
txt = {}
ops = defaultdict(int)
curf, cur, hdr = "", None, None
for row in csv.reader(io.StringIO(raw)):
    if not row:
        continue
    if row[0] in ("File Path", "File Name"):
        curf = row[1].split("/")[-1]
        continue
    if row[0] == "Line No":
        hdr = row
        ii = hdr.index("Instructions Executed")
        si = hdr.index("# Samples")
        continue
    if hdr is None or len(row) <= ii:
        continue
    if row[0]:
        cur = (curf, int(row[0]))
        txt[cur] = row[1]
    elif cur:
        try:
            n = int(row[ii])
        except ValueError:
            continue
        inst[cur] += n
        try:
            samp[cur] += int(row[si])
        except ValueError:
            pass
        tok = row[3].split()
        op = tok[1] if tok and tok[0].startswith("@") and len(tok) > 1 else (tok[0] if tok else "?")
        ops[op.split(".")[0]] += n
tot = sum(inst.values()) or 1
stot = sum(samp.values()) or 1
print(f"total warp instructions {tot}, stall samples {stot}")
print("-- by instructions executed")
for k, n in sorted(inst.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{k[0][:14]:14s}:{k[1]:4d} {100 * n / tot:5.1f}% inst {100 * samp[k] / stot:5.1f}% samp | {txt[k].strip()[:100]}")
print("-- by stall samples")
for k, n in sorted(samp.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{k[0][:14]:14s}:{k[1]:4d} {100 * inst[k] / tot:5.1f}% inst {100 * n / stot:5.1f}% samp | {txt[k].strip()[:100]}")
print("-- opcodes:", ", ".join(f"{k}:{100 * v / tot:.1f}%" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:24]))
