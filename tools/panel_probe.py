#!/usr/bin/env python
"""Where the time of a DAILY panel call goes after the kernels (C3b, device-resident raster): phases of
aggregate_dataset / aggregate_dataset_table, and a cProfile of the column assembly."""
import cProfile
import io
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import aggfly_b200 as af
    from aggfly_b200 import aggregate as agg, synthetic as syn
    wl = syn.make_workload("c3b_global_daily")
    dev = torch.device("cuda", 0)
    raster = wl.raster(dev, seed=1218)
    ds = wl.dataset(raster)
    w = wl.weights(ds)
    for name, fn in (("frame", af.aggregate_dataset), ("table", af.aggregate_dataset_table)):
        for i in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = fn(weights=w, dataset=ds, aggregator_dict=wl.spec)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            print(name, i, f"{dt * 1e3:.1f} ms", {k: round(v, 1) for k, v in agg.LAST_TRACE["phases_ms"].items()}, flush=True)
        n = len(out) if name == "frame" else out.num_rows
        print(name, "rows", n, flush=True)
    pr = cProfile.Profile()
    pr.enable()
    af.aggregate_dataset_table(weights=w, dataset=ds, aggregator_dict=wl.spec)
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18)
    print(s.getvalue())
    import tempfile
    d = tempfile.mkdtemp()
    t0 = time.perf_counter()
    af.write_table(out, os.path.join(d, "p.parquet"))
    print("write parquet", round((time.perf_counter() - t0) * 1e3, 1), "ms", os.path.getsize(os.path.join(d, "p.parquet")) / 1e6, "MB")


if __name__ == "__main__":
    main()
