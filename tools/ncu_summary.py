#!/usr/bin/env python3
"""Condense an ncu report (gpurun_out/prof_<w>.ncu-rep, scratch) into the small JSON kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_c2.ncu-rep profiles/r1_ncu_c2_v3_summary.json "workload text" "kernel text"
"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "sm__cycles_active.min",
        "sm__cycles_active.max", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size", "dram__bytes_read.sum",
        "dram__bytes_write.sum",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]


def main():
    rep, out, workload, kernel = sys.argv[1:5]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    metrics = {k: {"value": float(data[0][idx[k]].replace(",", "")), "unit": units[idx[k]]} for k in KEYS if k in idx}
    json.dump({"source": "ncu --set full --clock-control none --import-source on -k regex:simulate_kernel -s 4 -c 2 "
                         "(tools/gpu_profile.sh under gpurun, 1 x B200); first captured launch; raw report: " + rep +
                         " (scratch, not committed)",
               "kernel": kernel, "workload": workload, "metrics": metrics}, open(out, "w"), indent=1)
    for k, v in metrics.items():
        print("%-88s %14.6g %s" % (k, v["value"], v["unit"]))


if __name__ == "__main__":
    main()
