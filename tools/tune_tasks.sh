#!/bin/bash
# usage: tune_tasks.sh THREADS,MIN_CTAS,CHUNK,WINDOW,STREAMS ...   (task-graph solve kernel tuning)
cd "$(dirname "$0")/.."
last=""
for cfg in "$@"; do
  IFS=, read -r T C CH W S <<< "$cfg"
  if [ "$last" != "$T,$C" ]; then
    EA_NVCC_EXTRA="-DEA_TASK_THREADS=$T -DEA_TASK_MIN_CTAS=$C" python edge_alignment_b200/build.py --force > /dev/null 2>&1 || { echo "build failed $cfg"; continue; }
    last="$T,$C"
  fi
  EA_SOLVE_CHUNK=$CH EA_SOLVE_WINDOW=$W timeout -s KILL 120 python bench.py --steps 10 --warmup 3 --streams $S --cluster 0 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('cfg $cfg value %.0f ms/step %.3f solve_ms %.3f pre_ms %.3f frac %.3f'%(d['value'],d['ms_per_step'],r['kernel_ms_per_launch'],r['preprocess_ms_per_step'],r['frac']))"
done
python edge_alignment_b200/build.py --force > /dev/null 2>&1
