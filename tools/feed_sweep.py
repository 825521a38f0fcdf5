"""Pageable-host feed tuning: aggregate_dataset on a NumPy (pageable) global year for several staging
shapes.  usage (GPU box): python tools/feed_sweep.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import aggfly_b200 as af
from aggfly_b200 import stream, synthetic as syn

wl = syn.make_workload("c3_global_bins")
dev = torch.device("cuda", 0)
raster = wl.raster(dev, seed=1218)
host = raster.cpu().numpy()                      # pageable
ds_dev = wl.dataset(raster)
w = wl.weights(ds_dev)
del raster, ds_dev
torch.cuda.empty_cache()
hds = wl.dataset(host)
af.aggregate_dataset(weights=w, dataset=hds, aggregator_dict=wl.spec)     # warm-up (allocations, CSR)
for threads, slots, chunk_mb in [(4, 4, 256), (8, 8, 256), (8, 8, 128), (12, 12, 128), (16, 16, 64), (8, 8, 64), (12, 16, 256)]:
    stream.OPTIONS.update(staging_threads=threads, staging_slots=slots, chunk_bytes=chunk_mb << 20)
    ts = []
    for _ in range(2):
        t0 = time.perf_counter()
        df = af.aggregate_dataset(weights=w, dataset=hds, aggregator_dict=wl.spec)
        ts.append(time.perf_counter() - t0)
    print(f"threads {threads:2d} slots {slots:2d} chunk {chunk_mb:3d} MB: {min(ts) * 1e3:7.1f} ms  "
          f"({host.nbytes / min(ts) / 1e9:5.1f} GB/s end to end)  rows {len(df)}", flush=True)
