#!/bin/bash
# K1R iteration visit: regional tests, C3b timing next to the two-kernel path, ncu launch list + full capture.
# usage: tools/gpu_r2_k1r.sh <tag> [full]     (full: also the config-size parity + determinism tests)
set -u
TAG=${1:-r2k}
O=gpurun_out
mkdir -p $O
T="tests/test_gpu_regional.py"
[ "${2:-}" = "full" ] && T="tests/test_gpu_regional.py tests/test_gpu_config_parity.py"
timeout 900 python -m pytest $T -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$? t=$SECONDS"; tail -8 $O/${TAG}_pytest.log
timeout 300 python tools/regional_bench.py --steps 5 > $O/${TAG}_regional_c3b.jsonl 2> $O/${TAG}_regional_c3b.err; echo "regional rc=$? t=$SECONDS"; cat $O/${TAG}_regional_c3b.jsonl; tail -3 $O/${TAG}_regional_c3b.err
CMD="python tools/regional_bench.py --steps 2 --no-two"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:agf_ -c 40 --csv --log-file $O/${TAG}_launches_regional_c3b.csv $CMD > $O/${TAG}_ncu_launch.log 2>&1; echo "launch list rc=$? t=$SECONDS"
ncu --set full --clock-control none --import-source on -k regex:agf_k1_regional -s 2 -c 1 -o $O/${TAG}_prof_k1r -f $CMD > $O/${TAG}_ncu_full.log 2>&1; echo "full rc=$? t=$SECONDS"
