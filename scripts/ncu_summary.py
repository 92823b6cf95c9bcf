"""Selected columns of an `ncu --set full` report, one CSV row per captured launch (what profiles/*.csv hold).
usage: python scripts/ncu_summary.py report.ncu-rep > profiles/summary.csv"""
import csv
import subprocess
import sys

COLS = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keep = [c for c in COLS if c in hdr]
w = csv.writer(sys.stdout)
w.writerow(keep)
w.writerow([units[hdr.index(c)] for c in keep])
for r in rows[2:]:
    w.writerow([r[hdr.index(c)] for c in keep])
