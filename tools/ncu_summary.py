"""Summarise an .ncu-rep (ncu --set full) into the CSV layout used under profiles/: one row per captured launch with
the metrics the roofline discussion needs.   python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x.csv"""
import csv, io, subprocess, sys

COLS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed"]

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
head, units, body = rows[0], rows[1], rows[2:]
idx = [head.index(c) for c in COLS if c in head]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([head[i] for i in idx])
    w.writerow([units[i] for i in idx])
    for r in body:
        w.writerow([r[i] for i in idx])
print(f"{len(body)} launches -> {out}")
