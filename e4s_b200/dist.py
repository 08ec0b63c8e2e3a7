"""Multi-GPU driver of the synthesis path: independent faces are sharded across ranks (one process per GPU),
weights are replicated, and the only collective is the gather of the final images.

The reference has no multi-GPU inference path (its scripts are single `cuda:0`, swap_options.py:13); its only
collectives are DDP's in training (coach.py:46-85), which is out of scope.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of `n_items` faces owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_images(local: torch.Tensor, n_items: int = None) -> torch.Tensor:
    """All-gather per-rank image batches [b_r, 3, H, W] into [sum b_r, 3, H, W] on every rank (NCCL on GPU
    tensors, gloo on CPU tensors).  Ragged shards are padded to the largest shard for the collective."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    local = local.contiguous()
    if n_items is None:
        sizes = torch.tensor([local.shape[0]], device=local.device, dtype=torch.int64)
        all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
        dist.all_gather(all_sizes, sizes)
        counts = [int(s.item()) for s in all_sizes]
    else:
        counts = [shard_range(n_items, r, world)[1] - shard_range(n_items, r, world)[0] for r in range(world)]
    biggest = max(counts)
    if local.shape[0] < biggest:
        pad = local.new_zeros((biggest - local.shape[0],) + tuple(local.shape[1:]))
        local = torch.cat([local, pad], 0)
    out = local.new_empty((world * biggest,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, local)
    if all(c == biggest for c in counts):
        return out
    return torch.cat([out[r * biggest: r * biggest + counts[r]] for r in range(world)], 0)
