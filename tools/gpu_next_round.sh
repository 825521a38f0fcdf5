#!/bin/bash
# First GPU visit of the NEXT round: everything that was written after round 1's GPU budget was spent and is therefore
# opt-in, measured against the validated defaults.  Each step has its own timeout; outputs land in gpurun_out/<tag>_*.
# usage: tools/gpu_next_round.sh <tag>      (about 4-5 GPU-minutes)
set -u
TAG=${1:-r2a}
O=gpurun_out
mkdir -p $O
# 1. the opt-in paths through the whole GPU suite: device-side panel frame, inline chunk tables, multi-file feed
AGF_DEVICE_PANEL=1 AGF_TEST_INLINE_TABLES=1 AGF_TEST_UNVALIDATED=1 timeout 240 python -m pytest tests -m gpu -x -q \
    > $O/${TAG}_pytest_optin.log 2>&1; echo "pytest (opt-in paths) rc=$? t=$SECONDS"; tail -4 $O/${TAG}_pytest_optin.log
# 2. daily-panel workload end to end: literal panel assembly vs the device-side one
for dp in 0 1; do
  AGF_DEVICE_PANEL=$dp timeout 150 python bench.py --workload c3b_global_daily --no-cpu --e2e-steps 3 \
      > $O/${TAG}_bench_c3b_panel$dp.json 2> $O/${TAG}_bench_c3b_panel$dp.err
  echo "c3b e2e (device panel=$dp) rc=$? t=$SECONDS"; python - <<PY
import json
try:
    d = json.load(open("$O/${TAG}_bench_c3b_panel$dp.json")); e = d["e2e"]
    print("  e2e ms", e["step_ms"], "phases", e["phases_ms"][-1])
except Exception as exc:
    print("  (no e2e line:", exc, ")")
PY
done
# 3. small-chunk Blosc store: per-chunk tables from pageable memory vs riding the chunk's own copy
python - > $O/${TAG}_inline_tables.log 2>&1 <<'PY'
import json, subprocess, sys
for inline in (0, 1):
    code = ("import sys; sys.argv=['x','--only','time_major_24h_blosc','--reps','4'];"
            "from aggfly_b200 import stream; stream.OPTIONS['inline_chunk_tables']=bool(%d);"
            "import runpy; runpy.run_path('tools/zarr_feed_bench.py', run_name='__main__')" % inline)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=150)
    line = [l for l in out.stdout.splitlines() if l.startswith("{")]
    if line:
        d = json.loads(line[-1])["layouts"]
        print("inline_chunk_tables =", inline, {k: [round(x, 1) for x in v["ms"]] for k, v in d.items()})
    else:
        print("inline_chunk_tables =", inline, "FAILED", out.stderr[-400:])
PY
echo "inline tables rc=$? t=$SECONDS"; cat $O/${TAG}_inline_tables.log
# 4. ncu of the chunk kernels (one launch each) on the CONUS Blosc store
CMD="python tools/zarr_feed_bench.py --only reference_time_contiguous_blosc --reps 1"
timeout 90 $CMD > $O/${TAG}_zarr_plain.log 2>&1 &&
timeout 150 ncu --set full --clock-control none --import-source on -k regex:'agf_tile|agf_unshuffle|agf_copy_segments' -c 6 \
    -o $O/${TAG}_prof_chunk_kernels -f $CMD > $O/${TAG}_ncu_chunk.log 2>&1
echo "ncu chunk kernels rc=$? t=$SECONDS"
