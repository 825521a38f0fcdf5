#!/bin/bash
# Other chain shapes on the global hourly year, device-resident (which kernel each one lands on, ms per year), after the
# temporal parity suites.   usage: tools/gpu_r2_probe.sh <tag> [workload ...]
set -u
TAG=${1:-r3p}; shift
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_suite.py tests/test_gpu_config_parity.py -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$? t=$SECONDS"; tail -3 $O/${TAG}_pytest.log
for W in "$@"; do
  timeout 300 python bench.py --workload $W --no-e2e --no-cpu --no-extras --c4-years 0 --steps 3 > $O/${TAG}_$W.json 2> $O/${TAG}_$W.err; echo "$W rc=$? t=$SECONDS $(python tools/show_bench.py $O/${TAG}_$W.json 2>/dev/null | head -2 | tr '\n' ' ')"; tail -1 $O/${TAG}_$W.err | cut -c1-200
done
