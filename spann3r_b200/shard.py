"""Multi-GPU host logic of the inference path: independent sequences shard across ranks with no data-path
collective; the only exchanges are the timing reduction and an optional gather of results."""
from __future__ import annotations


def shard_indices(n_items: int, world_size: int, rank: int) -> list[int]:
    """Round-robin: sequence s runs on rank s % world_size (the reference's samplers slice per rank the same way,
    spann3r/datasets/__init__.py:27-39)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    return list(range(rank, n_items, world_size))


def max_over_ranks(value_ms: float, device=None) -> float:
    """Max of a per-rank scalar (CUDA-event milliseconds) over the process group; identity without one."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value_ms)
    t = torch.tensor([value_ms], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_objects(obj):
    """All ranks' small python objects on every rank (per-rank frame counts / timings)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out
