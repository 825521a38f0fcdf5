#!/bin/bash
# Decompression-engine visit: alignment probes (each in its own process: a failing engine operation poisons
# the CUDA context), then the Blosc tests and the zarr feed bench with the alignment the probes allow.
# usage: tools/gpu_de_round.sh <tag>
set -u
TAG=${1:-de}
O=gpurun_out
mkdir -p $O
for c in lz4_src_unaligned lz4_dst_unaligned; do
  timeout 60 python tools/probe_decompress.py --cases $c --mb 16 --out $O/${TAG}_probe_$c.json > $O/${TAG}_probe_$c.log 2>&1
  echo "probe $c rc=$?"; cat $O/${TAG}_probe_$c.json 2>/dev/null | cut -c1-600
done
ALIGN=0
grep -q '"round_trip_ok": true' $O/${TAG}_probe_lz4_src_unaligned.json 2>/dev/null || ALIGN=16
export AGF_DE_ALIGN=$ALIGN
echo "AGF_DE_ALIGN=$ALIGN"
timeout 150 python -m pytest tests/test_gpu_zarr.py -x -q -k "unshuffle or engine or blosc" > $O/${TAG}_pytest_de.log 2>&1
echo "pytest(de) rc=$?"; tail -8 $O/${TAG}_pytest_de.log
timeout 150 python tools/zarr_feed_bench.py --only blosc --out $O/${TAG}_zarr_feed_blosc.json > $O/${TAG}_zarr_blosc.log 2>&1
echo "zarr bench rc=$?"; tail -c 3000 $O/${TAG}_zarr_blosc.log
