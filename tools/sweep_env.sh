#!/bin/bash
# GPU-box helper: bench the default shape under different environment settings (no rebuild).
# usage: sweep_env.sh "VAR=val VAR2=val" ...     prints one line per setting; EA_* diagnostics from stderr follow it
cd "$(dirname "$0")/.."
for cfg in "$@"; do
  env $cfg timeout -s KILL 240 python bench.py --steps ${STEPS:-20} --warmup 3 --streams ${STREAMS:-592} --no-e2e --no-cpu-baseline 2>/tmp/sweep_err.txt | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; i=r['isolated']; g=r['gather_roof']
print('env [$cfg] value %.0f ms/step %.3f solve_live %.3f solve_alone %.3f pre_alone %.3f frac_live %.3f frac_alone %.3f roof_ms %.3f'%(d['value'],d['ms_per_step'],r['kernel_ms_per_launch'],i['kernel_ms_per_launch'],i['preprocess_ms_per_step'],r['frac'],i['frac'],g.get('launch_ms_at_roof',0)))" || echo "bench failed $cfg"
  grep "^\[EA_" /tmp/sweep_err.txt
done
