"""Clip sharding across the GPUs of one box (one process per GPU, torch.distributed for the plumbing).

Every clip is an independent sequence (the reference vmaps over the batch axis, seq_model.py:459-473, and
inference BatchNorm uses running statistics), so the path shards with NO data-path collective: each rank
runs the forward on a contiguous slice of the clip axis.  NCCL (gloo in CPU tests) is used only after the
forward, to gather outputs and reduce metrics.

Fixed-point caveat: ``compute_best`` exponents are global over the tensor a model
instance sees, so an integer result depends on its batch mates.  Sharding therefore gives *shard-local*
parity: each rank is bit-exact with the reference run on that rank's clips as its batch.
"""

from __future__ import annotations

from typing import Tuple


def shard_range(n_clips: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [start, stop) slice of the clip axis owned by ``rank`` (remainder goes to the low ranks)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    if n_clips < 0:
        raise ValueError("n_clips must be non-negative")
    base, rem = divmod(n_clips, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_outputs(local, n_clips: int, group=None):
    """All-gather per-rank outputs ``(b_local, L, F)`` back into clip order on every rank.

    Shards may be ragged (n_clips not divisible by the world size): ranks pad to the largest shard for the
    collective and the padding is dropped afterwards.
    """
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    sizes = [shard_range(n_clips, r, world)[1] - shard_range(n_clips, r, world)[0] for r in range(world)]
    mx = max(sizes)
    pad = local
    if local.shape[0] < mx:
        pad = torch.cat([local, local.new_zeros((mx - local.shape[0],) + tuple(local.shape[1:]))], 0)
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad.contiguous(), group=group)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], 0)


def reduce_metrics(values, op: str = "sum", group=None):
    """All-reduce a small 1-D tensor of metric accumulators (sums, counts, maxima)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return values
    ops = {"sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN}
    out = values.clone()
    dist.all_reduce(out, op=ops[op], group=group)
    return out
