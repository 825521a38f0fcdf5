"""One workload, a few hot-path steps -- the command profiled under ncu (see profiles/README.md)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from aggfly_b200 import engine, synthetic as syn
from aggfly_b200.aggregate import _device_csr, _plan

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c3_global_bins")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--stripes", type=int, default=0)
args = ap.parse_args()
dev = torch.device("cuda", 0)
wl = syn.make_workload(args.workload)
raster = wl.raster(dev, seed=1218)
ds = wl.dataset(raster)
w = wl.weights(ds)
names, stage = _plan(ds, wl.spec)
runner = engine.StageRunner(stage, wl.n_cells, dev, target_stripes=args.stripes)
csr = _device_csr(w, ds)
flat = raster.reshape(wl.n_time, wl.n_cells)
for _ in range(args.steps):
    ev = []
    res = runner.run(flat, k1_events=ev)
    panel = engine.run_spmm(csr, res)
    torch.cuda.synchronize()
    print("k1 ms:", [round(a.elapsed_time(b), 3) for a, b in ev], "panel nan rows:", int(torch.isnan(panel).any(-1).sum()))
