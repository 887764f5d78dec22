"""Multi-GPU sampling: the batch shards naturally (independent samples), one process per GPU.

No collective runs inside the DDIM loop; a single all-gather of the final latents (16 KiB/sample)
closes the job (north star: "batch-sharded across the 8 GPUs of one box with a single NCCL gather").
The global x_T / conditioning are drawn identically on every rank (same seed) and sliced, so the
union of the shards is bit-identical to the single-GPU run of the same global batch."""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of `total` units for `rank`; the first `total % world` ranks get one extra."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(t: Optional[torch.Tensor], rank: int, world: int) -> Optional[torch.Tensor]:
    if t is None:
        return None
    lo, hi = shard_bounds(t.shape[0], rank, world)
    return t[lo:hi].contiguous()


def gather_batch(local: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """All-gather shards of possibly unequal size back into the global batch order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_bounds(total, r, world) for r in range(world)]
    max_n = max(hi - lo for lo, hi in sizes)
    pad = local.new_zeros((max_n,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)], dim=0)


def sample_sharded(sample_fn: Callable[..., torch.Tensor], x_T: torch.Tensor,
                   cond: Optional[torch.Tensor] = None, uncond: Optional[torch.Tensor] = None,
                   gather: bool = True, group=None) -> torch.Tensor:
    """Run `sample_fn(x_T_shard, cond_shard, uncond_shard) -> latents` on this rank's shard of the
    global batch and (optionally) all-gather the latents."""
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    total = x_T.shape[0]
    out = sample_fn(shard(x_T, rank, world), shard(cond, rank, world), shard(uncond, rank, world))
    return gather_batch(out, total, group) if gather else out
