#!/usr/bin/env python3
"""Per-kernel summary (selected counters) of an ncu report, as CSV on stdout.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rX_summary.csv
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv", "--metrics", ",".join(METRICS)],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
keep = [i for i, h in enumerate(hdr) if h == "Kernel Name" or h in METRICS]
w = csv.writer(sys.stdout)
w.writerow([hdr[i] + (f" [{units[i]}]" if units[i] else "") for i in keep])
for r in rows[2:]:
    name = r[keep[0]]
    name = name.split("(")[0].replace("void ", "").replace("hpss::<unnamed>::", "").replace("unnamed>::", "")
    w.writerow([name] + [r[i] for i in keep[1:]])
