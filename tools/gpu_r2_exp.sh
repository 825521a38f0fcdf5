#!/bin/bash
# Timing experiments on C3b as they were run (profiles/r2_k1r_steps.jsonl): AGF_RG_EXP builds compiled parts of the kernel out
# (wrong results on purpose) and AGF_RG_SMEM_PAD lowered the occupancy; both switches were removed from the final sources.
# usage: tools/gpu_r2_exp.sh <tag> "<pads>" <variant> ...
set -u
TAG=$1; PADS=$2; shift 2
O=gpurun_out; mkdir -p $O
for P in $PADS; do
  AGF_RG_SMEM_PAD=$P timeout 300 python tools/regional_bench.py --steps 5 --no-two > $O/${TAG}_pad$P.jsonl 2> $O/${TAG}_pad$P.err; echo "pad=$P rc=$? t=$SECONDS $(python -c "import json,sys; d=json.loads(open('$O/${TAG}_pad$P.jsonl').readline()); print(round(d['ms'],3), d['ctas_per_sm'], d['smem'])")"; tail -2 $O/${TAG}_pad$P.err
done
for N in "$@"; do
  L=$PWD/aggfly_b200/csrc/variants/libaggfly_b200_$N.so
  AGF_B200_LIB=$L timeout 300 python tools/regional_bench.py --steps 5 --no-two > $O/${TAG}_var_$N.jsonl 2> $O/${TAG}_var_$N.err; echo "variant=$N rc=$? t=$SECONDS $(python -c "import json,sys; d=json.loads(open('$O/${TAG}_var_$N.jsonl').readline()); print(round(d['ms'],3))")"; tail -2 $O/${TAG}_var_$N.err
done
