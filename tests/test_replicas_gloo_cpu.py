"""world_size = 2 gloo test (CPU) of the multi-GPU bookkeeping used by bench.py: replicas exchange only the slowest
rank's time and the unit count; plus the row-shard helper of the tensor-parallel layout."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from onebit_b200 import replicas


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        agg = replicas.aggregate_throughput(local_units=128 * (rank + 1), local_ms=10.0 + 5.0 * rank)
        mx = replicas.max_over_ranks(3.0 - rank)
        dist.barrier()
        q.put((rank, agg, mx))
    finally:
        dist.destroy_process_group()


def test_replica_aggregation_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, agg, mx in out:
        assert agg["units"] == 128 + 256
        assert agg["ms"] == 15.0                      # slowest rank
        assert abs(agg["per_s"] - 384 / 0.015) < 1e-6
        assert mx == 3.0


def test_single_process_is_identity():
    assert replicas.max_over_ranks(1.5) == 1.5
    assert replicas.aggregate_throughput(10, 2.0)["per_s"] == 5000.0
