#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full) into the short text summary committed under profiles/.

  python profiles/summarize.py gpurun_out/prof.ncu-rep [updates_per_launch] > profiles/rNN_<kernel>.txt
"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def main():
    rep = sys.argv[1]
    upl = float(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")])
        vals = {}
        for k in KEYS:
            if k in hdr:
                vals[k] = r[hdr.index(k)]
                print(f"  {k:90s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
        if upl:
            def gb(x, u):
                x = float(x)
                return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
            rd = gb(vals["dram__bytes_read.sum"], units[hdr.index("dram__bytes_read.sum")])
            wr = gb(vals["dram__bytes_write.sum"], units[hdr.index("dram__bytes_write.sum")])
            inst = float(vals["smsp__inst_executed.sum"])
            print(f"  derived: updates/launch = {upl:.4g}; DRAM traffic = {rd + wr:.4g} B/launch = {(rd + wr) / upl:.3f} B/update "
                  f"(read {rd / upl:.3f}, write {wr / upl:.3f}); {inst * 32 / upl:.1f} lane-instructions/update")
        print()


if __name__ == "__main__":
    main()
