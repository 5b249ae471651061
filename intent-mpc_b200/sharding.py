"""Batch-index sharding of independent QPs across the GPUs of one box: contiguous ranges
[g*B/G, (g+1)*B/G) per rank, one process / engine per GPU, no collective on the solve path; one all_gather of the
solutions afterwards, only so that one rank can verify them (NCCL over NVLink with CUDA tensors, gloo in the CPU tests).
In the receding-horizon configuration the sharding unit is the scenario (all intent candidates of a scenario stay on
one GPU): pass `unit` = candidates per scenario."""
from __future__ import annotations

import numpy as np


def shard_bounds(B: int, world: int, unit: int = 1):
    """[(lo, hi)] per rank; ranges are contiguous, cover [0, B) exactly and cut only at multiples of `unit`."""
    if B % unit:
        raise ValueError("B must be a multiple of unit")
    nu = B // unit
    cuts = [(nu * g) // world * unit for g in range(world + 1)]
    return [(cuts[g], cuts[g + 1]) for g in range(world)]


def shard(mb, rank: int, world: int, unit: int = 1):
    lo, hi = shard_bounds(mb.B, world, unit)[rank]
    return mb.slice(lo, hi), (lo, hi)


def gather_rows(local, B: int, world: int, unit: int = 1, group=None, device=None):
    """all_gather of per-rank row blocks of unequal length into the full [B, ...] array (same order as the unsharded
    batch).  `local` is a numpy array or torch tensor holding this rank's rows."""
    import torch
    import torch.distributed as dist
    bounds = shard_bounds(B, world, unit)
    t = torch.as_tensor(local)
    if device is not None:
        t = t.to(device)
    width = max(hi - lo for lo, hi in bounds)
    pad = torch.zeros((width,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([parts[g][: hi - lo] for g, (lo, hi) in enumerate(bounds)], dim=0)
