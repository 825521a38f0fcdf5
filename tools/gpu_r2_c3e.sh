#!/bin/bash
# Chains of the reference tests' shape on the global grid (bins / date -> sum / year, alone and next to the polynomial of
# the daily mean), hourly bins by year, the daily panel through both paths: the whole GPU suite, then the workloads timed
# (AGF_BINS_BY_EDGES=0: the bin-by-bin count of the same build).   usage: tools/gpu_r2_c3e.sh <tag>
set -u
TAG=${1:-r3m}
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$? t=$SECONDS"; tail -3 $O/${TAG}_pytest.log
for E in 2 0; do
for W in c3e_global_bins_date_year c3d_global_hourly_bins; do
  AGF_BINS_BY_EDGES=$E timeout 300 python bench.py --workload $W --no-e2e --no-cpu --no-extras --c4-years 0 --steps 5 > $O/${TAG}_${W}_e$E.json 2> $O/${TAG}_${W}_e$E.err; echo "$W edges=$E rc=$? t=$SECONDS"; python tools/show_bench.py $O/${TAG}_${W}_e$E.json 2>/dev/null | head -2; tail -2 $O/${TAG}_${W}_e$E.err
done
done
timeout 300 python tools/regional_bench.py --steps 5 2>/dev/null | cut -c1-110
