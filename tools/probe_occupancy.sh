#!/bin/bash
# GPU-box helper: gather-roof probe at 4 / 2 / 1 resident CTAs of 256 threads per SM (32 / 16 / 8 warps)
cd "$(dirname "$0")/.."
for KB in 0 100 200; do
EA_PROBE_SMEM_KB=$KB python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); g=d['roofline']['gather_roof']
print('probe smem $KB KB: rates', ['%.1f G/s'%(r/1e9) for r in g['point_gathers_per_s_by_level']], 'launch_ms_at_roof %.2f'%g['launch_ms_at_roof'])"
done
