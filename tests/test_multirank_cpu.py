"""world_size-2 gloo tests (CPU) of the N>1 host logic: rank partitioning, timing aggregation and the reference arm's
rank-0-only behaviour.  The data path itself has no collective for the pair shard."""
import os
import subprocess
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from edge_alignment_b200 import sharding


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 4096, 44458, 2_000_000):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 4096 independent pairs (BASELINE configs[2]) sharded in contiguous blocks
        ref = np.arange(4096); now = np.arange(4096) + 10000
        r, n = sharding.shard_pairs(ref, now, rank, world)
        mine = torch.zeros(4096, dtype=torch.int64); mine[r] = 1
        dist.all_reduce(mine)
        covered = bool((mine == 1).all())
        # timing: slowest rank defines the step, units add up
        value, ms = sharding.aggregate_throughput(len(r), 10.0 * (rank + 1), dist)
        q.put((rank, covered, len(r), value, ms))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_partition_and_timing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    out = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for rank, covered, n, value, ms in out:
        assert covered and n == 2048
        assert ms == 20.0 and abs(value - 4096 / 0.020) < 1e-6


def test_reference_arm_other_ranks_exit_quietly():
    """bench.py --impl reference under torchrun: only rank 0 works and prints; other ranks exit 0 without output."""
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == ""
