#!/bin/bash
# Scaling visit on an N-GPU box: the bench under torchrun at each N (ours, then the CPU arm at the largest N).
# usage: tools/gpu_scale.sh <tag> "<N list>"        e.g.  gpurun --gpus 8 -- 'bash tools/gpu_scale.sh r1s "1 2 4 8"'
set -u
TAG=${1:-r1s}
NS=${2:-"1 2"}
O=gpurun_out
mkdir -p $O
free -g | head -2; nproc; nvidia-smi -L
PORT=29520
for n in $NS; do
  PORT=$((PORT + 1))
  if [ "$n" = "1" ]; then
    python bench.py --gpus 1 --steps 10 --warmup 3 > $O/${TAG}_bench_n$n.json 2> $O/${TAG}_bench_n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $PORT \
        bench.py --gpus $n --steps 10 --warmup 3 > $O/${TAG}_bench_n$n.json 2> $O/${TAG}_bench_n$n.err
  fi
  echo "N=$n rc=$?"; cat $O/${TAG}_bench_n$n.json; tail -2 $O/${TAG}_bench_n$n.err
done
