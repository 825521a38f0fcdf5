#!/bin/bash
# Round 2 measurement hygiene: ncu --set full of the temporal kernel on C1 and C5, and of the chunk kernels
# (tile placement / transposition, unshuffle, segment copy, linear unpack) as the zarr and packed feeds launch them.
set -u
TAG=${1:-r2h}
O=gpurun_out
mkdir -p $O
for wl in c1_conus_tavg c5_cmip_gdd; do
  CMD="python bench.py --workload $wl --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras"
  $CMD > $O/${TAG}_plain_$wl.json 2> $O/${TAG}_plain_$wl.err; echo "$wl plain rc=$? t=$SECONDS"
  ncu --set full --clock-control none --import-source on -k regex:agf_k1 -s 3 -c 1 -o $O/${TAG}_prof_k1_$wl -f $CMD > $O/${TAG}_ncu_$wl.log 2>&1; echo "$wl ncu rc=$? t=$SECONDS"
done
CMD="python tools/zarr_feed_bench.py --reps 1"
$CMD > $O/${TAG}_zarr_feed_plain.json 2> $O/${TAG}_zarr_feed_plain.err; echo "zarr feed plain rc=$? t=$SECONDS"
ncu --set full --clock-control none -k regex:"agf_tile|agf_unshuffle|agf_copy_segments" -c 12 -o $O/${TAG}_prof_chunk -f $CMD > $O/${TAG}_ncu_chunk.log 2>&1; echo "chunk ncu rc=$? t=$SECONDS"
cat > /tmp/packed_probe.py <<'P'
import numpy as np, pandas as pd, torch, sys
sys.path.insert(0, ".")
import aggfly_b200 as af
from aggfly_b200 import engine
q = torch.from_numpy(np.random.default_rng(0).integers(-30000, 30000, (480, 721, 1440), dtype=np.int16)).pin_memory()
p = af.PackedRaster(q, 0.002, 280.0, -32767.0)
for _ in range(2):
    r = engine.to_device(p)
torch.cuda.synchronize()
print(r.shape, float(r[0, 0, 0]))
P
python /tmp/packed_probe.py > $O/${TAG}_packed_probe.log 2>&1; echo "packed probe rc=$? t=$SECONDS"
ncu --set full --clock-control none -k regex:agf_tile_linear -s 2 -c 2 -o $O/${TAG}_prof_linear -f python /tmp/packed_probe.py > $O/${TAG}_ncu_linear.log 2>&1; echo "linear ncu rc=$? t=$SECONDS"
