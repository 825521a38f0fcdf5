#!/bin/bash
# K1R iteration visit (round 2, late): regional tests, then C3b timings of the edge-counting variants side by side
# (AGF_BINS_BY_EDGES = 1: float32 compares, 2: bfloat16-packed compares), an ncu launch list and one full capture.
# usage: tools/gpu_r2_k1r2.sh <tag> [full] [nocap]
set -u
TAG=${1:-r2k}
O=gpurun_out
mkdir -p $O
T="tests/test_gpu_regional.py"
[ "${2:-}" = "full" ] && T="tests/test_gpu_regional.py tests/test_gpu_config_parity.py"
timeout 900 python -m pytest $T -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$? t=$SECONDS"; tail -8 $O/${TAG}_pytest.log
for V in 1 2; do
  AGF_BINS_BY_EDGES=$V timeout 300 python tools/regional_bench.py --steps 5 $([ $V = 1 ] && echo --no-two) > $O/${TAG}_regional_c3b_v$V.jsonl 2> $O/${TAG}_regional_c3b_v$V.err; echo "regional v$V rc=$? t=$SECONDS"; cut -c1-400 $O/${TAG}_regional_c3b_v$V.jsonl; tail -3 $O/${TAG}_regional_c3b_v$V.err
done
timeout 300 python tools/regional_bench.py --steps 5 --no-two --noise 1.0 > $O/${TAG}_regional_c3b_noise1.jsonl 2>&1; cut -c1-200 $O/${TAG}_regional_c3b_noise1.jsonl
[ "${3:-}" = "nocap" ] && exit 0
CMD="python tools/regional_bench.py --steps 2 --no-two"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:agf_ -c 40 --csv --log-file $O/${TAG}_launches_regional_c3b.csv $CMD > $O/${TAG}_ncu_launch.log 2>&1; echo "launch list rc=$? t=$SECONDS"
ncu --set full --clock-control none --import-source on -k regex:agf_k1_regional -s 2 -c 1 -o $O/${TAG}_prof_k1r -f $CMD > $O/${TAG}_ncu_full.log 2>&1; echo "full rc=$? t=$SECONDS"
