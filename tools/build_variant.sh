#!/bin/bash
# Builds the library from the current sources with extra nvcc flags into aggfly_b200/csrc/variants/libaggfly_b200_<name>.so
# (AGF_B200_LIB selects it at run time).  usage: tools/build_variant.sh <name> "<flags>"
set -eu
NAME=$1; FLAGS=${2:-}; EXTRA=${3:-}   # EXTRA: more units to rebuild (names without .cu)
ROOT=$(cd "$(dirname "$0")/.." && pwd)
B=/tmp/agf_variant_$NAME
rm -rf $B && mkdir -p $B/aggfly_b200 $B/include
cp $ROOT/include/*.h $B/include/
mkdir -p $B/aggfly_b200/csrc && cp $ROOT/aggfly_b200/csrc/*.cu $ROOT/aggfly_b200/csrc/*.cuh $ROOT/aggfly_b200/csrc/*.h $ROOT/aggfly_b200/csrc/Makefile $B/aggfly_b200/csrc/
# objects that do not include agf_regional.cuh are reused
for o in agf_api agf_geom agf_tile agf_k1_f32_tma_single agf_k1_f32_tma_two agf_k1_f32_tma_uni agf_k1_f32_ldg agf_k1_f64_tma agf_k1_f64_ldg; do
  cp $ROOT/aggfly_b200/csrc/$o.o $B/aggfly_b200/csrc/ 2>/dev/null || true
done
for u in $EXTRA; do rm -f $B/aggfly_b200/csrc/$u.o; done
(cd $B/aggfly_b200/csrc && touch -d '2 hours ago' *.cu *.cuh *.h && touch agf_k1_f32_regional.cu agf_rplan.cu && make -j8 NVCCFLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xptxas -v -I../../include --fmad=false -ccbin /usr/bin/g++ $FLAGS" > build.log 2>&1 || (tail -20 build.log; exit 1))
mkdir -p $ROOT/aggfly_b200/csrc/variants
cp $B/aggfly_b200/csrc/libaggfly_b200.so $ROOT/aggfly_b200/csrc/variants/libaggfly_b200_$NAME.so
grep -E "registers|spill" $B/aggfly_b200/csrc/agf_k1_f32_regional.ptxas.log | paste - - | grep -B0 "72 reg\|spill" | head -3
echo "built variants/libaggfly_b200_$NAME.so"
