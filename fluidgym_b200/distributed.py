"""Environment-batch sharding across GPUs: environments are independent, so each
rank owns a contiguous chunk of the batch and there is NO collective on the data path.  The only
collectives are the bookkeeping ones of a benchmark / learner (max of timings, sum of step counts)."""
from __future__ import annotations


def shard_envs(total_envs: int, world_size: int, rank: int) -> range:
    """Contiguous chunk of ``range(total_envs)`` owned by ``rank`` (sizes differ by at most one,
    like ``torch.chunk`` used by the reference's ParallelFluidEnv.step, envs/parallel_env.py:233-240)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(total_envs, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def _reduce(x: float, op: str, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return float(t.item())


def max_over_ranks(x: float, device=None) -> float:
    return _reduce(x, "max", device)


def sum_over_ranks(x: float, device=None) -> float:
    return _reduce(x, "sum", device)
