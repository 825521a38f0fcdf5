#!/bin/bash
# Round 2, first visit: the whole GPU suite incl. the config-size oracle parity + determinism tests, then the default bench.
set -u
TAG=${1:-r2a}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc; free -g | head -2
timeout 900 python -m pytest tests -m gpu -x -q --durations=15 > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$? t=$SECONDS"; tail -25 $O/${TAG}_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench_c3.err; echo "bench rc=$? t=$SECONDS"
head -c 1500 $O/${TAG}_bench_c3.json
