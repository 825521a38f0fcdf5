#!/bin/bash
# Last round-2 visit: smoke, the whole GPU suite, C3b through the one-kernel path with its ncu launch list and full captures,
# both bench arms, the ncu launch list of the bench command.   usage: tools/gpu_r2_final.sh <tag>
set -u
TAG=${1:-r2final}
O=gpurun_out
mkdir -p $O
timeout 200 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$? t=$SECONDS"; tail -1 $O/${TAG}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$? t=$SECONDS"; tail -9 $O/${TAG}_pytest.log
timeout 300 python tools/regional_bench.py --steps 5 > $O/${TAG}_regional_c3b.jsonl 2> $O/${TAG}_regional_c3b.err; echo "regional rc=$? t=$SECONDS"; cut -c1-160 $O/${TAG}_regional_c3b.jsonl; grep -o '"max_rel_vs_two.*' $O/${TAG}_regional_c3b.jsonl; tail -2 $O/${TAG}_regional_c3b.err
for S in 1.0 0.3; do timeout 300 python tools/regional_bench.py --steps 5 --no-two --noise $S 2>/dev/null | cut -c1-100 | tee -a $O/${TAG}_regional_noise.jsonl; done
CMD="python tools/regional_bench.py --steps 2 --no-two"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:agf_ -c 40 --csv --log-file $O/${TAG}_launches_regional_c3b.csv $CMD > $O/${TAG}_ncu_launch.log 2>&1; echo "launch list rc=$? t=$SECONDS"
ncu --set full --clock-control none --import-source on -k regex:agf_k1_regional -s 2 -c 1 -o $O/${TAG}_prof_k1r -f $CMD > $O/${TAG}_ncu_full.log 2>&1; echo "full rc=$? t=$SECONDS"
ncu --set full --clock-control none --import-source on -k regex:agf_regional_merge -s 2 -c 1 -o $O/${TAG}_prof_k1rm -f $CMD > $O/${TAG}_ncu_full_m.log 2>&1; echo "full merge rc=$? t=$SECONDS"
timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$? t=$SECONDS"; tail -3 $O/${TAG}_bench.err
python tools/show_bench.py $O/${TAG}_bench.json | cut -c1-400
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; echo "ref rc=$? t=$SECONDS"; cut -c1-300 $O/${TAG}_bench_ref.json
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:agf_ -c 80 --csv --log-file $O/${TAG}_launches_bench.csv $CMD > $O/${TAG}_ncu_launch_bench.log 2>&1; echo "bench launch list rc=$? t=$SECONDS"
