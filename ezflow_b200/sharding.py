"""Batch sharding of image pairs over the GPUs of one node.

The similarity path has no cross-pair dependency (reference: a batched bmm, pairwise.py:82-85; DataParallel scatter on dim 0,
engine/eval.py:306), so pairs are sharded along the batch dimension, one process per GPU, with NO data-path collective.
torch.distributed is only used for the barrier around the timed region and for max-over-ranks of the per-rank timings."""

import torch


def shard_range(n_items, world_size, rank):
    """Contiguous, balanced block [lo, hi) of `n_items` owned by `rank` (sizes differ by at most one)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside [0, {world_size})")
    base, rem = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(values, device=None):
    """Element-wise maximum of a list of floats over all ranks (identity when torch.distributed is not initialised)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(v) for v in values]
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.tolist()]


def sum_over_ranks(values, device=None):
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(v) for v in values]
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(v) for v in t.tolist()]
