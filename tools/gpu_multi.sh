#!/bin/bash
# GPU-box helper for gpurun --gpus N: the multi-rank GPU tests and the sharded / scaled bench lines.
# usage: gpu_multi.sh <N> <tag> [workloads...]     (workloads: config5 config3 config2)
cd "$(dirname "$0")/.."
N=$1; TAG=$2; shift 2
O=gpurun_out; mkdir -p $O
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ -z "$SKIP_TESTS" ]; then timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s > $O/${TAG}_multi_tests.log 2>&1; echo "multi tests rc=$?"; tail -4 $O/${TAG}_multi_tests.log; fi
for w in "$@"; do
  if [ "$w" = "config2" ]; then timeout 900 $RUN bench.py --gpus $N --steps 30 --warmup 5 > $O/${TAG}_${w}_${N}gpu.json 2> $O/${TAG}_${w}_${N}gpu.err
  else timeout 900 $RUN bench.py --gpus $N --workload $w > $O/${TAG}_${w}_${N}gpu.json 2> $O/${TAG}_${w}_${N}gpu.err; fi
  echo "$w x$N rc=$?"; tail -c 600 $O/${TAG}_${w}_${N}gpu.json; echo
done
