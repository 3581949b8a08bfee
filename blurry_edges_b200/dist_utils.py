"""The two collectives of the path, kept separate so that the CPU test-suite can run them under gloo."""
from __future__ import annotations

import torch
import torch.distributed as dist


def sync_loss_normalisers(mask_count: torch.Tensor, local_patches: int, group=None):
    """Depth-term mask count (global_training.py:127 divides by the count of the WHOLE batch) and patch count of the global
    batch.  mask_count: int64[1] of this rank, summed in place over `group`.  Returns (mask_count, global_patches)."""
    if group is None and not (dist.is_available() and dist.is_initialized()):
        return mask_count, local_patches
    world = dist.get_world_size(group)
    if world > 1:
        dist.all_reduce(mask_count, op=dist.ReduceOp.SUM, group=group)
    return mask_count, local_patches * world


def reduce_accumulator(acc: torch.Tensor, group=None, dst_group_rank: int = 0):
    """Sum the ranks' partial fold accumulators ([1,H,W,16]) onto one rank (big-image block sharding)."""
    if dist.get_world_size(group) > 1:
        dst = dist.get_global_rank(group, dst_group_rank) if group is not None else dst_group_rank
        dist.reduce(acc, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return acc
