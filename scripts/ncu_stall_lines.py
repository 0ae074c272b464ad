"""Top source lines of an ncu report for one stall reason.

    python scripts/ncu_stall_lines.py rep.ncu-rep stall_no_inst [n]
"""
import csv, io, subprocess, sys
from collections import defaultdict
rep, reason = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cnt = defaultdict(int); txt = {}; opc = defaultdict(int)
curf, cur, hdr = "", None, None
for row in csv.reader(io.StringIO(raw)):
    if not row: continue
    if row[0] in ("File Path", "File Name"): curf = row[1].split("/")[-1]; continue
    if row[0] == "Line No": hdr = row; ri = hdr.index(reason); continue
    if hdr is None or len(row) <= ri: continue
    if row[0]: cur = (curf, int(row[0])); txt[cur] = row[1]
    elif cur:
        try: n = int(row[ri])
        except ValueError: continue
        cnt[cur] += n
        tok = row[3].split()
        op = tok[1] if tok and tok[0].startswith("@") and len(tok) > 1 else (tok[0] if tok else "?")
        opc[op.split(".")[0]] += n
tot = sum(cnt.values()) or 1
print(reason, "samples", tot)
for k, n in sorted(cnt.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{k[0][:14]:14s}:{k[1]:4d} {100*n/tot:5.1f}% | {txt[k].strip()[:110]}")
print("opcodes:", ", ".join(f"{k}:{100*v/tot:.1f}%" for k, v in sorted(opc.items(), key=lambda kv: -kv[1])[:12]))
