#!/bin/bash
# GPU-box helper: chamfer DT kernel variants (warp kernel vs block kernel with P pixels per thread)
cd "$(dirname "$0")/.."
echo "warp kernel:"; bash tools/sweep_cluster.sh 592 0
for P in 4 8 12 16; do echo "block kernel P=$P:"; EA_DT_BLOCK=1 EA_DT_P=$P bash tools/sweep_cluster.sh 592 0; done
