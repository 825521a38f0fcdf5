#!/usr/bin/env python
"""Top SASS instructions of a kernel by stall samples, and sample shares of code regions split at barriers.
usage: tools/ncu_source_top.py source.csv [n]   (source.csv = `ncu -i rep --page source --csv`)"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = rows[1]
ix = {k: i for i, k in enumerate(h)}
body = rows[2:]
tot = sum(int(r[ix["# Samples"]]) for r in body)
tot_inst = sum(int(r[ix["Instructions Executed"]]) for r in body)
print(f"total samples {tot}, warp instructions {tot_inst}")
stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
print("-- regions (split at BAR / mbarrier waits): start_line, first instr, samples %, inst %")
start, acc_s, acc_i = 0, 0, 0
for i, r in enumerate(body):
    acc_s += int(r[ix["# Samples"]]); acc_i += int(r[ix["Instructions Executed"]])
    if "BAR.SYNC" in r[ix["Source"]] or i == len(body) - 1:
        print(f"  lines {start:5d}-{i:5d}  samples {100*acc_s/tot:5.1f}%  inst {100*acc_i/tot_inst:5.1f}%   ends at {r[ix['Source']].strip()[:40]}")
        start, acc_s, acc_i = i + 1, 0, 0
print("-- top instructions by samples")
order = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]]))[:n]
for i in order:
    r = body[i]
    top = sorted(((int(r[ix[s]]), s) for s in stalls), reverse=True)[:2]
    print(f"  line {i:5d} {100*int(r[ix['# Samples']])/tot:5.2f}%  exec {int(r[ix['Instructions Executed']]):>10d}  {r[ix['Source']].strip()[:60]:60s} {top}")
