#!/bin/bash
# Round 2 visit 3: the new host-feed paths (ring of device windows, packed int16) + the extended bench line.
set -u
TAG=${1:-r2n}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_regional.py tests/test_gpu_zarr.py -m gpu -x -q -k "ring or packed or zarr or streamed" > $O/${TAG}_pytest_new.log 2>&1; echo "pytest rc=$? t=$SECONDS"; tail -15 $O/${TAG}_pytest_new.log
timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$? t=$SECONDS"; tail -5 $O/${TAG}_bench.err
python tools/show_bench.py $O/${TAG}_bench.json
