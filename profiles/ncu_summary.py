"""Summarise an .ncu-rep (raw page) into a small table: python profiles/ncu_summary.py file.ncu-rep"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
M = [
    ("time_ms", "gpu__time_duration.sum"), ("dram_rd_GB", "dram__bytes_read.sum"), ("dram_wr_GB", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("regs", "launch__registers_per_thread"), ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("fp64_pipe_pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    ("alu_pipe_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    ("fma_pipe_pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    ("inst_M", "smsp__inst_executed.sum"), ("l1_hit_pct", "l1tex__t_sector_hit_rate.pct"),
    ("l2_hit_pct", "lts__t_sector_hit_rate.pct"), ("lts_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("stall_long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("stall_wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
    ("stall_short_sb", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
    ("stall_math_throttle", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
    ("stall_barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
    ("stall_branch", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"),
    ("stall_not_selected", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"),
    ("stall_lg_throttle", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"),
    ("local_ld_sectors", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum"),
]
ki = h.index("Kernel Name")
names = [r[ki].split("(")[0][-28:] for r in rows[2:]]
print("%-22s" % "metric", *["%14s" % n[:14] for n in names])
for label, m in M:
    if m not in h:
        continue
    i = h.index(m)
    vals = []
    for r in rows[2:]:
        try:
            v = float(r[i].replace(",", ""))
            if label == "inst_M":
                v /= 1e6
            vals.append("%14.3f" % v)
        except ValueError:
            vals.append("%14s" % r[i][:14])
    print("%-22s" % label, *vals)
