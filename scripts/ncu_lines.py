"""Aggregate 'Instructions Executed' per CUDA source line from an ncu report (needs -lineinfo).

    python scripts/ncu_lines.py gpurun_out/prof.ncu-rep [top_n]
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    per_line = defaultdict(lambda: [0, 0, ""])  # (file, line) -> [inst, samples, text]
    cur_file, cur = "", None
    hdr = None
    for row in csv.reader(io.StringIO(raw)):
        if not row:
            continue
        if row[0] == "File Path":
            cur_file = row[1].split("/")[-1]
            continue
        if row[0] == "Line No":
            hdr = row
            i_inst = hdr.index("Instructions Executed")
            i_samp = hdr.index("# Samples")
            continue
        if hdr is None or len(row) <= i_inst:
            continue
        if row[0]:  # a CUDA source line
            cur = (cur_file, int(row[0]))
            per_line[cur][2] = row[1].strip()
        elif cur is not None:  # SASS instruction belonging to the current line
            try:
                per_line[cur][0] += int(row[i_inst])
                per_line[cur][1] += int(row[i_samp])
            except ValueError:
                pass
    total = sum(v[0] for v in per_line.values()) or 1
    tsamp = sum(v[1] for v in per_line.values()) or 1
    print(f"total warp instructions {total}, samples {tsamp}")
    for (f, ln), (inst, samp, text) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100.0 * inst / total:5.1f}% inst {100.0 * samp / tsamp:5.1f}% stall  {f}:{ln:<4d} {text[:100]}")


if __name__ == "__main__":
    main()
