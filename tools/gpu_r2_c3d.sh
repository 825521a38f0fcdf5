#!/bin/bash
# Hourly bins by year (c3d): parity tests of the temporal kernels, then the workload timed on the default build and on the
# named variants.   usage: tools/gpu_r2_c3d.sh <tag> [variant ...]
set -u
TAG=$1; shift
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_suite.py tests/test_gpu_config_parity.py -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$? t=$SECONDS"; tail -3 $O/${TAG}_pytest.log
B="python bench.py --workload c3d_global_hourly_bins --no-e2e --no-cpu --no-extras --c4-years 0 --steps 5"
timeout 300 $B > $O/${TAG}_c3d.json 2> $O/${TAG}_c3d.err; echo "c3d rc=$? t=$SECONDS"; python tools/show_bench.py $O/${TAG}_c3d.json | head -3; tail -2 $O/${TAG}_c3d.err
for N in "$@"; do
  AGF_B200_LIB=$PWD/aggfly_b200/csrc/variants/libaggfly_b200_$N.so timeout 300 $B > $O/${TAG}_c3d_$N.json 2> $O/${TAG}_c3d_$N.err; echo "c3d[$N] rc=$? t=$SECONDS"; python tools/show_bench.py $O/${TAG}_c3d_$N.json | head -3; tail -2 $O/${TAG}_c3d_$N.err
done
