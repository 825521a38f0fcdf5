#!/bin/bash
# One GPU-box visit: parity tests, the bench line (both arms), the other workloads, and (NCU=1) the
# ncu evidence of the same bench command (launch list + one full capture of the temporal kernel).
# usage: [NCU=1] [WL="c1_conus_tavg ..."] [NCU_WL=c3_global_bins] tools/gpu_round.sh <tag>
set -u
TAG=${1:-r1}
O=gpurun_out
NCU=${NCU:-0}
WL=${WL:-"c1_conus_tavg c2_conus_gdd c3b_global_daily c5_cmip_gdd"}
NCU_WL=${NCU_WL:-c3_global_bins}
mkdir -p $O
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/${TAG}_smoke.log
python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $O/${TAG}_pytest.log
python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; cat $O/${TAG}_bench.json; tail -3 $O/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; echo "ref rc=$?"; cat $O/${TAG}_bench_ref.json
for wl in $WL; do
  python bench.py --workload $wl --no-e2e --no-cpu > $O/${TAG}_bench_$wl.json 2> $O/${TAG}_bench_$wl.err; echo "$wl rc=$?"; cat $O/${TAG}_bench_$wl.json; tail -3 $O/${TAG}_bench_$wl.err
done
if [ "$NCU" = "1" ]; then
  CMD="python bench.py --workload $NCU_WL --steps 2 --warmup 3 --no-e2e --no-cpu"
  $CMD > $O/${TAG}_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:agf_ -c 60 --csv \
      --log-file $O/${TAG}_launches.csv $CMD > $O/${TAG}_ncu_launch.log 2>&1
  echo "launch list rc=$?"
  $CMD > $O/${TAG}_plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:agf_k1 -s 3 -c 1 \
      -o $O/${TAG}_prof_k1 -f $CMD > $O/${TAG}_ncu_full.log 2>&1
  echo "full capture (k1) rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:agf_spmm -s 3 -c 1 \
      -o $O/${TAG}_prof_k2 -f $CMD > $O/${TAG}_ncu_full_k2.log 2>&1
  echo "full capture (k2) rc=$?"
fi
