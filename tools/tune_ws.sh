#!/bin/bash
# GPU-box helper: warp-specialised kernel shapes.  usage: tune_ws.sh MATH,PROD,MATH_REGS,GATHER_REGS ...
cd "$(dirname "$0")/.."
for cfg in "$@"; do
  IFS=, read -r M P MR GR BO <<< "$cfg"
  EA_NVCC_EXTRA="-DEA_WS_MATH=$M -DEA_WS_PROD=$P -DEA_WS_MATH_REGS=$MR -DEA_WS_GATHER_REGS=$GR -DEA_WS_BACKOFF_NS=${BO:-0}" python edge_alignment_b200/build.py --force > /dev/null 2>&1 || { echo "build failed $cfg"; continue; }
  timeout -s KILL 150 python bench.py --steps 10 --warmup 3 --cluster -3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; i=r['isolated']
print('ws $cfg value %.0f ms/step %.3f solve_live %.3f solve_alone %.3f'%(d['value'],d['ms_per_step'],r['kernel_ms_per_launch'],i['kernel_ms_per_launch']))"
done
python edge_alignment_b200/build.py --force > /dev/null 2>&1
