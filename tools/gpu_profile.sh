#!/bin/bash
# GPU-box helper: tests, default bench line, ncu launch list and one full capture of the solve kernel.
# usage: gpu_profile.sh <tag>        outputs under gpurun_out/<tag>_*
cd "$(dirname "$0")/.."
TAG=${1:-rX}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/${TAG}_tests.log
timeout 600 python bench.py > $O/${TAG}_bench_default.json 2> $O/${TAG}_bench_default.err; echo "bench rc=$?"
CMD="python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > $O/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^(k_|ea_k)" -c 600 --csv --log-file $O/${TAG}_launches_ncu.csv $CMD > $O/${TAG}_ncu_list.log 2>&1
echo "ncu list rc=$?"
$CMD > $O/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ea_k_solve_batch -s 5 -c 1 -f -o $O/${TAG}_solve_full $CMD > $O/${TAG}_ncu_full.log 2>&1
echo "ncu full rc=$?"
ncu -i $O/${TAG}_solve_full.ncu-rep --page raw --csv > $O/${TAG}_solve_raw.csv 2>/dev/null
ncu -i $O/${TAG}_solve_full.ncu-rep --page source --csv > $O/${TAG}_solve_src.csv 2>/dev/null
EA_SOLVE_DEBUG=1 STEPS=20 bash tools/sweep_solve.sh "512,1" > $O/${TAG}_sweep.log 2>&1; cat $O/${TAG}_sweep.log
