#!/bin/bash
# GPU-box helper: bench the solve kernels against each other.  usage: sweep_cluster.sh [streams] kernel...
cd "$(dirname "$0")/.."
S=${1:-592}; shift
for K in "$@"; do
python bench.py --steps 10 --warmup 3 --streams $S --cluster $K --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('streams $S kernel $K value %.0f ms/step %.3f solve_ms %.3f pre_ms %.3f Gpe/s %.2f'%(d['value'],d['ms_per_step'],r['kernel_ms_per_launch'],r['preprocess_ms_per_step'],d['config']['point_evals_per_s']/1e9))"
done
