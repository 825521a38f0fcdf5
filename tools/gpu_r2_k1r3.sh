#!/bin/bash
# K1R visit: regional + config-size parity tests, C3b timing (default build, then the named variants), ncu launch list and
# full captures of the scan and merge kernels.   usage: tools/gpu_r2_k1r3.sh <tag> [variant ...]
set -u
TAG=$1; shift
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_regional.py tests/test_gpu_config_parity.py -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$? t=$SECONDS"; tail -4 $O/${TAG}_pytest.log
timeout 300 python tools/regional_bench.py --steps 5 > $O/${TAG}_regional_c3b.jsonl 2> $O/${TAG}_regional_c3b.err; echo "regional rc=$? t=$SECONDS"; cut -c1-120 $O/${TAG}_regional_c3b.jsonl; grep -o '"max_rel_vs_two.*' $O/${TAG}_regional_c3b.jsonl; tail -2 $O/${TAG}_regional_c3b.err
for N in "$@"; do
  L=$PWD/aggfly_b200/csrc/variants/libaggfly_b200_$N.so
  AGF_B200_LIB=$L timeout 300 python tools/regional_bench.py --steps 5 --no-two > $O/${TAG}_var_$N.jsonl 2> $O/${TAG}_var_$N.err; echo "variant=$N rc=$? t=$SECONDS $(cut -c1-110 $O/${TAG}_var_$N.jsonl)"; tail -2 $O/${TAG}_var_$N.err
done
CMD="python tools/regional_bench.py --steps 2 --no-two"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:agf_ -c 40 --csv --log-file $O/${TAG}_launches_regional_c3b.csv $CMD > $O/${TAG}_ncu_launch.log 2>&1; echo "launch list rc=$? t=$SECONDS"; grep -E "agf_k1_regional|agf_regional_merge" $O/${TAG}_launches_regional_c3b.csv | tail -2 | cut -c1-60,200-
ncu --set full --clock-control none --import-source on -k regex:agf_k1_regional -s 2 -c 1 -o $O/${TAG}_prof_k1r -f $CMD > $O/${TAG}_ncu_full.log 2>&1; echo "full rc=$? t=$SECONDS"
ncu --set full --clock-control none --import-source on -k regex:agf_regional_merge -s 2 -c 1 -o $O/${TAG}_prof_k1rm -f $CMD > $O/${TAG}_ncu_full_m.log 2>&1; echo "full merge rc=$? t=$SECONDS"
