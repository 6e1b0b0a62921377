#!/bin/bash
# GPU-box helper: rebuild with extra -D flags and bench each.  usage: tune_defs.sh "<flags>" "<flags>" ...
cd "$(dirname "$0")/.."
for cfg in "$@"; do
  EA_NVCC_EXTRA="$cfg" python edge_alignment_b200/build.py --force > /dev/null 2>&1 || { echo "build failed $cfg"; continue; }
  python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; i=r['isolated']
print('cfg [$cfg] value %.0f ms/step %.3f solve_live %.3f solve_alone %.3f pre_alone %.3f'%(d['value'],d['ms_per_step'],r['kernel_ms_per_launch'],i['kernel_ms_per_launch'],i['preprocess_ms_per_step']))"
done
python edge_alignment_b200/build.py --force > /dev/null 2>&1
