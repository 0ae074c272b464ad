"""Turn the raw files of scripts/gpu_evidence.sh (gpurun_out/) into the tracked summaries under profiles/.

    python scripts/make_profile_summary.py r01a

Writes profiles/<tag>_launches.csv (the ncu launch list of `bench.py --steps 2 --warmup 3`, one row per
launch), profiles/<tag>_summary.md (per-kernel shares of the step, the bench line, and for each
`ncu --set full` capture the metrics, stall reasons and hottest source lines).  No GPU needed.
"""
import csv
import io
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)
out = []

# ---- bench line ----------------------------------------------------------------------------------
bpath = os.path.join(G, f"bench_{tag}.json")
if os.path.exists(bpath) and os.path.getsize(bpath):
    line = json.loads(open(bpath).read().strip().splitlines()[-1])
    out.append(f"# Profile summary {tag}\n")
    out.append("## bench.py line (same box, not under ncu)\n")
    out.append("```json\n" + json.dumps(line, indent=1) + "\n```\n")
    with open(os.path.join(P, f"{tag}_bench.json"), "w") as f:
        f.write(json.dumps(line) + "\n")

# ---- launch list ---------------------------------------------------------------------------------
lpath = os.path.join(G, f"launches_{tag}.csv")
if os.path.exists(lpath):
    rows = [r for r in csv.reader(open(lpath)) if r and not r[0].startswith("==")]
    hdr, rows = rows[0], rows[1:]
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    with open(os.path.join(P, f"{tag}_launches.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel", "grid", "block", "duration_ns"])
        for r in rows:
            w.writerow([r[0], r[ki].split("(")[0][:70], r[gi], r[hdr.index("Block Size")], r[vi]])
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = r[ki].split("(")[0]
        if name.startswith("void "):
            name = name[5:]
        if "k_idf" in name or "k_vg_prefix" in name or "k_init_positions" in name:
            continue  # world generation, its prefix sums and the reset are set-up, not part of the step
        agg[name][0] += 1
        agg[name][1] += float(r[vi].replace(",", ""))
    tot = sum(v[1] for v in agg.values()) or 1.0
    out.append("## ncu launch list of `bench.py --steps 4 --warmup 3 --step-only` (the bench's step loop alone)\n")
    out.append("Per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes. "
               "Set-up launches (IDF generation, `k_vg_prefix` -- 4.8 ms once per world set for 1024 environments --, the reset) are left out of the shares.\n")
    out.append("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {n} | {t / 1e3:.1f} | {100 * t / tot:.1f}% |")
    out.append("")

# ---- full captures -------------------------------------------------------------------------------
KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
]
traffic = {}
# kernel-name fragment of the capture -> key of bench.py's roofline.kernels
TRAFFIC_KEYS = {"k_vg_reduce<": "A/B k_vg_reduce(streaming)", "k_vg_partition<1, 1>": "k_vg_partition<moments>",
                "k_vg_partition<(int)1, (bool)1>": "k_vg_partition<moments>", "k_voronoi_window": "k_voronoi_window",
                "k_obs_fused<1": "k_obs_fused<window Voronoi>", "k_obs_fused<(bool)1": "k_obs_fused<window Voronoi>",
                "k_obs_fused<0": "k_obs_fused(local+obstacle+comm)", "k_obs_fused<(bool)0": "k_obs_fused(local+obstacle+comm)",
                "k_mark_windows": "k_mark_windows"}
PROFILE_ENVS = int(os.environ.get("PROFILE_ENVS", 256))  # scripts/gpu_evidence.sh profiles with this many environments


def _num(x):
    return float(x.replace(",", "") or 0)


def _bytes(value, unit):
    return _num(value) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


for fn in sorted(os.listdir(G)):
    if not (fn.startswith(f"prof_{tag}_") and fn.endswith(".ncu-rep")):
        continue
    rep = os.path.join(G, fn)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rws = list(csv.reader(io.StringIO(raw)))
    if len(rws) < 3:
        continue
    hdr, units = rws[0], rws[1]
    for r in rws[2:]:
        out.append(f"## ncu --set full: `{r[hdr.index('Kernel Name')][:80]}` ({fn})\n")
        for short, key in TRAFFIC_KEYS.items():
            if short in r[hdr.index("Kernel Name")]:
                ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
                total = _bytes(r[ir], units[ir]) + _bytes(r[iw], units[iw])
                traffic[key] = {"dram_bytes_per_launch": total, "envs": PROFILE_ENVS, "dram_bytes_per_env": total / PROFILE_ENVS,
                                "duration_us_under_ncu": _num(r[hdr.index("gpu__time_duration.sum")]), "report": fn}
        out.append("| metric | value | unit |\n|---|---:|---|")
        for k in KEYS:
            if k in hdr:
                out.append(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |")
        st = [(float(r[i].replace(",", "") or 0), h) for i, h in enumerate(hdr)
              if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
        out.append("\nStall reasons (warps per issue-active cycle): " + ", ".join(
            f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}"
            for v, h in sorted(st, reverse=True)[:8]) + "\n")
    top = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_toplines.py"), rep, "12"],
                         capture_output=True, text=True).stdout
    out.append("```\n" + "\n".join(l[:170] for l in top.splitlines()) + "\n```\n")

if traffic:
    with open(os.path.join(P, "traffic.json"), "w") as f:
        json.dump({"source": f"ncu --set full captures of scripts/gpu_evidence.sh {tag} (profiles/{tag}_summary.md)",
                   "kernels": traffic}, f, indent=1)
with open(os.path.join(P, f"{tag}_summary.md"), "w") as f:
    f.write("\n".join(out) + "\n")
print("wrote", os.path.join(P, f"{tag}_summary.md"))
