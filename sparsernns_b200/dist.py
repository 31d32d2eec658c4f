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


def parse_cpulist(text: str) -> list:
    """'0-3,8,10-11' -> [0, 1, 2, 3, 8, 10, 11] (the format of /sys/.../local_cpulist)."""
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (best effort; returns what was done).

    Pinned host buffers are placed by first touch, so a rank that runs on its GPU's node gets host memory on that
    node and its H2D / D2H copies do not cross the socket interconnect.  Matters for the end-to-end number when
    8 ranks stream 2 x 50 GB/s each; has no effect on the device-resident numbers.
    """
    import os

    import torch

    info = {"bound": False}
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        with open(f"{base}/numa_node") as f:
            info["numa_node"] = int(f.read().strip())
        with open(f"{base}/local_cpulist") as f:
            cpus = parse_cpulist(f.read())
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info.update(bound=True, n_cpus=len(allowed))
    except Exception as e:  # no sysfs entry (virtualised PCI), no permission, ...: run unbound
        info["error"] = f"{type(e).__name__}: {e}"
    return info
