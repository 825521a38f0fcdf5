#!/usr/bin/env python
"""Prints the interesting keys of a bench.py JSON line.  usage: tools/show_bench.py file.json"""
import json
import sys

d = json.load(open(sys.argv[1]))
print("value %.4g %s  ms_per_step %.3f  n_gpus %s" % (d["value"], d["unit"], d["ms_per_step"], d["n_gpus"]))
r = d.get("roofline") or {}
print("roofline frac %.3f  k1_ms %.3f  achieved %.0f GB/s" % (r.get("frac", 0), r.get("k1_ms", 0), r.get("achieved", 0)))
for k in ("e2e", "e2e_resident", "e2e_packed"):
    e = d.get(k)
    if e:
        print(k, {x: e[x] for x in ("value", "ms_per_step", "step_ms", "h2d_bytes_per_step", "error", "device_raster_equals_numpy_decode", "phases_ms") if x in e},
              (e.get("feed") or {}).get("h2d_gbs"))
    else:
        print(k, e)
print("c4", d.get("c4"))
print("c4_packed", d.get("c4_packed"))
for w in d.get("workloads") or []:
    if "error" in w:
        print("  ", w)
    else:
        print("   %-26s %8.3f ms  k1 %8.3f ms  frac %.3f  %s" % (w["workload"], w["ms_per_step"], w["k1_ms"], w["roofline"]["frac"], w["path"][:40]))
print("cpu", d.get("cpu_baseline"))
print("clocks", d.get("clocks"))
