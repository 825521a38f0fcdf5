#!/bin/bash
# Chains of the reference tests' shape on the global grid (bins / date -> sum / year, alone and next to the polynomial of
# the daily mean), hourly bins by year: the whole GPU suite, then the three workloads timed.   usage: tools/gpu_r2_c3e.sh <tag>
set -u
TAG=${1:-r3m}
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$? t=$SECONDS"; tail -3 $O/${TAG}_pytest.log
for W in c3e_global_bins_date_year c3f_global_bins_date_year_only c3d_global_hourly_bins; do
  timeout 300 python bench.py --workload $W --no-e2e --no-cpu --no-extras --c4-years 0 --steps 5 > $O/${TAG}_$W.json 2> $O/${TAG}_$W.err; echo "$W rc=$? t=$SECONDS"; python tools/show_bench.py $O/${TAG}_$W.json 2>/dev/null | head -2; tail -2 $O/${TAG}_$W.err
done
