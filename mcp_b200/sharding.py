"""Partitioning of a θ batch over GPUs/ranks.  Instances are independent, so the only
multi-GPU logic is: contiguous column blocks, no data-path collective, timing = max over ranks."""
from __future__ import annotations

from typing import Tuple


def shard_range(B: int, rank: int, world: int) -> Tuple[int, int]:
    """Columns [begin, end) of rank `rank` out of `world`: the same split libmcpb200 uses across
    devices inside one process (`make_shards` in csrc/mcpb200.cpp)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return B * rank // world, B * (rank + 1) // world


def reduce_max_ms(ms: float, group=None) -> float:
    """Max over ranks of a device-measured duration (torch.distributed, any backend)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(ms)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([float(ms)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def reduce_sum_int(v: int, group=None) -> int:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return int(v)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([int(v)], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t.item())
