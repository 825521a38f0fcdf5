#!/bin/bash
# One visit, several builds of the library side by side (tools/build_variant.sh): regional tests on the default build,
# C3b timings for the values of an environment knob (AGF_RG_BLOCKS: the period-block experiment, since removed) and for
# every variant named on the command line.
# usage: tools/gpu_r2_variants.sh <tag> "<blocks list>" <variant> [<variant> ...]     (a variant suffixed with +t also runs the tests)
set -u
TAG=$1; BLOCKS=$2; shift 2
O=gpurun_out
mkdir -p $O
T="tests/test_gpu_regional.py"
timeout 900 python -m pytest $T -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$? t=$SECONDS"; tail -4 $O/${TAG}_pytest.log
for B in $BLOCKS; do
  AGF_RG_BLOCKS=$B timeout 300 python tools/regional_bench.py --steps 5 --no-two > $O/${TAG}_blocks$B.jsonl 2> $O/${TAG}_blocks$B.err; echo "blocks=$B rc=$? t=$SECONDS $(cut -c1-110 $O/${TAG}_blocks$B.jsonl)"; tail -2 $O/${TAG}_blocks$B.err
done
for V in "$@"; do
  N=${V%+t}
  L=$PWD/aggfly_b200/csrc/variants/libaggfly_b200_$N.so
  if [ "$V" != "$N" ]; then
    AGF_B200_LIB=$L timeout 600 python -m pytest $T -m gpu -x -q > $O/${TAG}_pytest_$N.log 2>&1; echo "pytest[$N] rc=$? t=$SECONDS"; tail -3 $O/${TAG}_pytest_$N.log
  fi
  AGF_B200_LIB=$L timeout 300 python tools/regional_bench.py --steps 5 > $O/${TAG}_var_$N.jsonl 2> $O/${TAG}_var_$N.err; echo "variant=$N rc=$? t=$SECONDS"; cut -c1-110 $O/${TAG}_var_$N.jsonl; grep -o '"max_rel_vs_two.*' $O/${TAG}_var_$N.jsonl; tail -2 $O/${TAG}_var_$N.err
done
