#!/usr/bin/env python
"""Condense an .ncu-rep with several kernel launches (--set full) into one table, a column per launch.
usage: ncu_summary_multi.py <report.ncu-rep> ["header text"] > profiles/<name>.txt"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2:]
want = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_requests.sum", "lts__t_sectors.sum", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
ix = {h: i for i, h in enumerate(hdr)}
print("# " + (sys.argv[2] if len(sys.argv) > 2 else rep))
names = [v[ix["Kernel Name"]].split("(")[0] for v in vals]
print("%-86s" % "Kernel Name" + " | ".join("%16s" % n[:16] for n in names))
for k in want:
    if k not in ix:
        continue
    cells = []
    for v in vals:
        x = v[ix[k]].replace(",", "")
        try:
            f = float(x)
            x = "%d" % f if f == int(f) and abs(f) < 1e15 else "%.4g" % f
        except ValueError:
            pass
        cells.append("%16s" % x[:16])
    print("%-86s" % k + " | ".join(cells) + " " + units[ix[k]])
