#!/bin/bash
# Last GPU visit of the round: smoke, the GPU tests, both bench arms; then, while time remains, the global
# zarr check on a latitude band and the ncu launch list of the bench command.
# usage: tools/gpu_final.sh <tag>
set -u
TAG=${1:-final}
O=gpurun_out
mkdir -p $O
timeout 60 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$? t=$SECONDS"; tail -1 $O/${TAG}_smoke.log
timeout 120 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$? t=$SECONDS"; tail -4 $O/${TAG}_pytest.log
timeout 120 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$? t=$SECONDS"; cut -c1-900 $O/${TAG}_bench.json; tail -2 $O/${TAG}_bench.err
timeout 60 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; echo "ref rc=$? t=$SECONDS"; cut -c1-300 $O/${TAG}_bench_ref.json
if [ $SECONDS -lt 150 ]; then
  timeout 80 python tools/zarr_global_bench.py --rows 240 --reps 2 --out $O/${TAG}_zarr_band.json > $O/${TAG}_zarr_band.log 2>&1; echo "zarr band rc=$? t=$SECONDS"; tail -c 1800 $O/${TAG}_zarr_band.log
fi
if [ $SECONDS -lt 185 ]; then
  CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
  timeout 50 $CMD > $O/${TAG}_plain.log 2>&1 &&
  timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:agf_ -c 60 --csv \
      --log-file $O/${TAG}_launches.csv $CMD > $O/${TAG}_ncu_launch.log 2>&1
  echo "launch list rc=$? t=$SECONDS"
fi
