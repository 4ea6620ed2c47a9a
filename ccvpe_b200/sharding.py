"""Batch sharding of independent query/aerial pairs across the GPUs of one box.

Pair b's outputs depend only on pair b in eval mode (reference models.py:196 multiplies the ground map of sample b with
the aerial window of sample b; BN uses running statistics), so the path shards over the batch with NO data-path
collective.  The only communication is an optional gather of the tiny decoded poses (<= 40 bytes per pair) for
reporting; it works with any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced shard [lo, hi) of `n_items` for `rank`: the first n_items % world ranks get one extra."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world: %d/%d" % (rank, world))
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(grd: torch.Tensor, sat: torch.Tensor, rank: int, world: int):
    lo, hi = shard_bounds(grd.shape[0], rank, world)
    return grd[lo:hi], sat[lo:hi]


def gather_poses(pose: Dict[str, torch.Tensor], n_total: int, group=None) -> Optional[Dict[str, torch.Tensor]]:
    """All-gathers the per-shard pose tensors (idx, rc, cs, angle, valid) back into batch order on every rank.
    Shards may be ragged (n_total not divisible by world): every rank pads to the largest shard."""
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized():
        return pose
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bounds = [shard_bounds(n_total, r, world) for r in range(world)]
    longest = max(hi - lo for lo, hi in bounds)
    out: Dict[str, torch.Tensor] = {}
    for key, t in pose.items():
        pad_shape = (longest,) + tuple(t.shape[1:])
        padded = torch.zeros(pad_shape, dtype=t.dtype, device=t.device)
        padded[: t.shape[0]] = t
        parts: List[torch.Tensor] = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(parts, padded, group=group)
        out[key] = torch.cat([parts[r][: bounds[r][1] - bounds[r][0]] for r in range(world)], dim=0)
    return out
