#!/usr/bin/env python
"""Print a compact view of a gpurun call log written by tools/gpu_round.sh."""
import json
import sys

for line in open(sys.argv[1]).read().splitlines():
    if line.startswith("{"):
        try:
            d = json.loads(line)
        except Exception:
            print(line[:300])
            continue
        r = d.get("roofline") or {}
        e = d.get("e2e") or {}
        print(d.get("impl", "ours"), d["config"]["workload"], "N", d.get("n_gpus"), "ms/step", round(d["ms_per_step"], 3),
              "value %.3e" % d["value"], "| k1_ms", round(r.get("k1_ms", 0), 3), "GB/s", round(r.get("achieved", 0)),
              "| e2e", "%.3e" % e["value"] if e else None, round(e.get("ms_per_step", 0), 1) if e else "", e.get("feed"),
              "| cpu", d.get("cpu_baseline") and "%.3e" % d["cpu_baseline"]["value"], "| clocks", d.get("clocks"))
    elif not line.startswith("[gpurun] merged"):
        print(line[:300])
