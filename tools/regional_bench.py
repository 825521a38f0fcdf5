#!/usr/bin/env python
"""Times the one-kernel temporal + regional path on a workload (default: C3b, the global daily panel) for a few
(periods per unit, ring blocks) settings, next to the two-kernel path.  Prints one JSON line per setting."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3b_global_daily")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--no-two", action="store_true")
    ap.add_argument("--noise", type=float, default=3.0, help="sigma of the field's iid hourly noise (bench default 3.0)")
    args = ap.parse_args()
    import torch
    from aggfly_b200 import engine, synthetic as syn
    from aggfly_b200.aggregate import _device_csr, _plan
    dev = torch.device("cuda", 0)
    wl = syn.make_workload(args.workload)
    raster = wl.raster(dev, seed=1218, noise_sigma=args.noise)
    ds = wl.dataset(raster)
    w = wl.weights(ds)
    csr = _device_csr(w, ds)
    names, stage = _plan(ds, wl.spec)
    flat = raster.reshape(wl.n_time, wl.n_cells)
    n_lat, n_lon = len(wl.grid.latitude), len(wl.grid.longitude)

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps

    ref = None
    if not args.no_two:
        two = engine.StageRunner(stage, wl.n_cells, dev)
        ms = timed(lambda: engine.run_spmm(csr, two.run(flat)))
        ref = engine.run_spmm(csr, two.run(flat)).clone()
        torch.cuda.synchronize()
        print(json.dumps({"path": "two kernels (K1 -> X -> K2)", "ms": ms}), flush=True)
        two.close()
        del two
        torch.cuda.empty_cache()
    rr = engine.RegionalRunner(stage, csr, n_lat, n_lon)
    if not rr.supported:
        print(json.dumps({"path": "regional", "supported": False}))
        return
    ms = timed(lambda: rr.run(flat))
    out = {"path": "one kernel (+ merge)", "noise_sigma": args.noise, "ms": ms, "workspace_GB": rr.info.workspace_bytes / 1e9,
           "lps": int(rr.info.lanes_per_slot), "ctas_per_sm": int(rr.info.ctas_per_sm), "smem": int(rr.info.smem_bytes),
           "active_tiles": int(rr.plan.info.n_active_tiles), "tiles": int(rr.plan.info.n_tiles), "slots": int(rr.plan.info.n_slots),
           "partial_rows": int(rr.plan.info.n_partial_rows), "max_slots_per_tile": int(rr.plan.info.max_slots_per_tile),
           "read_GB": rr.algorithmic_input_bytes() / 1e9, "GBps": rr.algorithmic_input_bytes() / ms / 1e6}
    if ref is not None:
        p = rr.run(flat).panel
        torch.cuda.synchronize()
        ok = ~torch.isnan(ref)
        out["nan_equal"] = bool(torch.equal(torch.isnan(p), torch.isnan(ref)))
        out["max_rel_vs_two"] = float(((p[ok] - ref[ok]).abs() / ref[ok].abs().clamp_min(1e-2)).max())
        out["frac_bit_equal"] = float((p[ok] == ref[ok]).double().mean())
    print(json.dumps(out), flush=True)
    rr.close()


if __name__ == "__main__":
    main()
