#!/bin/bash
# GPU-box helper: rebuild the solve kernel with different shapes and bench each (tuning experiments).
# usage: sweep_solve.sh "THREADS,UNROLL[,extra -D flags]" ...     prints one line per configuration
cd "$(dirname "$0")/.."
for cfg in "$@"; do
  IFS=, read -r T U X <<< "$cfg"
  EA_NVCC_EXTRA="-DEA_SOLVE_THREADS=$T -DEA_EVAL_UNROLL=$U $X" python edge_alignment_b200/build.py --force > /dev/null 2>&1 || { echo "build failed $cfg"; continue; }
  timeout -s KILL 240 python bench.py --steps ${STEPS:-20} --warmup 3 --streams ${STREAMS:-592} --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; i=r['isolated']; g=r['gather_roof']
print('cfg [$cfg] value %.0f ms/step %.3f solve_live %.3f solve_alone %.3f pre_alone %.3f frac_live %.3f frac_alone %.3f roof_ms %.3f'%(d['value'],d['ms_per_step'],r['kernel_ms_per_launch'],i['kernel_ms_per_launch'],i['preprocess_ms_per_step'],r['frac'],i['frac'],g.get('launch_ms_at_roof',0)))" || echo "bench failed $cfg"
done
python edge_alignment_b200/build.py --force > /dev/null 2>&1
