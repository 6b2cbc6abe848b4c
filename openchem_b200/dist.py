"""Multi-GPU plumbing of the data-parallel path: one process per GPU, molecules sharded across
ranks, no collective inside the hot path; the only exchange is DDP's gradient all-reduce, exactly as
run.py:65-70, 234-236 sets it up. Kept tiny on purpose: torch.distributed is plumbing, not the product."""
from __future__ import annotations

import os


def env_world():
    """(world_size, rank, local_rank) from the launcher environment (launch.py:185-207 / torchrun)."""
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_seed(base, rank):
    """Every rank draws its own molecules: a per-rank seed shards the synthetic batch without any exchange."""
    return int(base) + int(rank)


def shard_bounds(n_items, world, rank):
    """Contiguous shard [lo, hi) of n_items for `rank` (sizes differ by at most one; DistributedSampler-like)."""
    base, rem = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(value, device=None):
    """Max of a scalar over all ranks (timings are reported as the slowest rank's)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def wrap_ddp(model, local_rank=None):
    """DistributedDataParallel exactly as run.py:235-236 (default buckets, broadcast_buffers=True: rank 0's BatchNorm
    running statistics are re-broadcast every forward; BatchNorm itself stays per-rank, it is NOT SyncBN)."""
    import torch
    if local_rank is None:
        return torch.nn.parallel.DistributedDataParallel(model)
    return torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank], output_device=local_rank)
