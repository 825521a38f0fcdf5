#!/usr/bin/env python
"""Key metrics of every kernel in an .ncu-rep, as CSV (run where ncu is installed; no GPU needed).
usage: tools/ncu_summary.py report.ncu-rep [out.csv]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
out = io.StringIO()
w = csv.writer(out)
w.writerow(["kernel", "metric", "value", "unit"])
for row in rows[2:]:
    d = dict(zip(hdr, row))
    u = dict(zip(hdr, units))
    for k in KEYS:
        if k in d:
            w.writerow([d.get("Kernel Name", "")[:90], k, d[k], u.get(k, "")])
text = out.getvalue()
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text)
else:
    print(text)
