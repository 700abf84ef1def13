"""Replica-parallel bookkeeping for multi-GPU runs (`bench.py --gpus N`): the decode path shards per model replica,
so ranks exchange nothing on the data path; they only agree on the slowest rank's time and the total work."""
from __future__ import annotations

import torch
import torch.distributed as dist


def max_over_ranks(value: float, device="cpu", group=None) -> float:
    """Elapsed time of the slowest rank (device-timed value in, same value on every rank out)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def sum_over_ranks(value: float, device="cpu", group=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t.item())


def aggregate_throughput(local_units: float, local_ms: float, device="cpu", group=None) -> dict:
    """Whole-job throughput of independent replicas: all units processed / slowest rank's time."""
    total = sum_over_ranks(local_units, device, group)
    ms = max_over_ranks(local_ms, device, group)
    return {"units": total, "ms": ms, "per_s": total / (ms * 1e-3)}
