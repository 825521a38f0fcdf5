#!/bin/bash
# Round 2 visit 4: whole GPU suite (NetCDF-4, ring, packed, K1R v7), C3b timing + ncu, the bench line.
set -u
TAG=${1:-r2s}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$? t=$SECONDS"; tail -14 $O/${TAG}_pytest.log
timeout 300 python tools/regional_bench.py --steps 5 > $O/${TAG}_regional_c3b.jsonl 2> $O/${TAG}_regional_c3b.err; echo "regional rc=$? t=$SECONDS"; cat $O/${TAG}_regional_c3b.jsonl; tail -3 $O/${TAG}_regional_c3b.err
CMD="python tools/regional_bench.py --steps 2 --no-two"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:agf_ -c 40 --csv --log-file $O/${TAG}_launches_regional_c3b.csv $CMD > $O/${TAG}_ncu_launch.log 2>&1; echo "launch list rc=$? t=$SECONDS"
ncu --set full --clock-control none --import-source on -k regex:agf_k1_regional -s 2 -c 1 -o $O/${TAG}_prof_k1r -f $CMD > $O/${TAG}_ncu_full.log 2>&1; echo "full rc=$? t=$SECONDS"
ncu --set full --clock-control none --import-source on -k regex:agf_regional_merge -s 2 -c 1 -o $O/${TAG}_prof_k1rm -f $CMD > $O/${TAG}_ncu_full_m.log 2>&1; echo "full merge rc=$? t=$SECONDS"
timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$? t=$SECONDS"; tail -3 $O/${TAG}_bench.err
python tools/show_bench.py $O/${TAG}_bench.json
