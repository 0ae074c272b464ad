"""Print the metrics that matter from an .ncu-rep (run where ncu is installed, no GPU needed).

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [--stalls]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_xu.sum",
    "sm__inst_executed_pipe_lsu.sum", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__cycles_active.avg",
    "local_load", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("===", r[hdr.index("Kernel Name")][:60], "id", r[0])
        for k in KEYS:
            if k in hdr:
                print(f"  {k:70s} {r[hdr.index(k)]} {units[hdr.index(k)]}")
        if "--stalls" in sys.argv:
            st = [(float(r[i].replace(",", "") or 0), h) for i, h in enumerate(hdr)
                  if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
            for v, h in sorted(st, reverse=True)[:8]:
                print(f"  stall {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):30s} {v:.2f}")


if __name__ == "__main__":
    main()
