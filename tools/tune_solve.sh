#!/bin/bash
# GPU-box helper: rebuild the solve kernel with different launch shapes and bench each (tuning experiments)
# usage: tune_solve.sh THREADS,MIN_CTAS,FLUSH,STREAMS,CLUSTER ...
cd "$(dirname "$0")/.."
for cfg in "$@"; do
  IFS=, read -r T C F S K <<< "$cfg"
  EA_NVCC_EXTRA="-DEA_SOLVE_THREADS=$T -DEA_SOLVE_MIN_CTAS=$C -DEA_FLUSH_EVERY=$F" python edge_alignment_b200/build.py --force > /dev/null 2>&1 || { echo "build failed $cfg"; continue; }
  python bench.py --steps 20 --warmup 3 --streams $S --cluster ${K:-0} --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; i=r['isolated']
print('cfg $cfg value %.0f ms/step %.3f solve_live %.3f solve_alone %.3f pre_alone %.3f'%(d['value'],d['ms_per_step'],r['kernel_ms_per_launch'],i['kernel_ms_per_launch'],i['preprocess_ms_per_step']))"
done
python edge_alignment_b200/build.py --force > /dev/null 2>&1
