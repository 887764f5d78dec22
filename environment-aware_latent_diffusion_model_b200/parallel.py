"""Multi-GPU sampling and training: the batch shards naturally (independent samples), one process per GPU.

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


# ---- training: data-parallel gradient all-reduce ---------------------------------------------------------
class GradBuckets:
    """Data-parallel gradient exchange for the hand-written backward (reference: pytorch-lightning DDP,
    main.py:741-745 / SURVEY.md section 8e): the only collective of a training step.

    All parameter gradients live in ONE flat fp32 buffer (`param.grad` are views into it), ordered by the time
    the backward finishes them: output head, output blocks (last to first), middle block, input blocks (last to
    first), then the tensors that are only complete at the very end (timestep-embedding MLP, every ResBlock's
    emb_layers and conv1 bias, the cross-attention K/V projections of the context).  `UNetTrainEngine.backward`
    reports every finished UNet block through `block_done`, which launches an asynchronous all-reduce (NCCL:
    on its own stream, overlapping the rest of the backward) for every bucket that is now complete.
    `finish()` launches the remainder, waits, and the gradients are the mean over ranks."""

    def __init__(self, unet, bucket_mb: float = 64.0, group=None):
        self.group = group
        self.unet = unet
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        late_keys = ("time_embed.", ".emb_layers.", ".in_layers.2.bias", ".attn2.to_k.", ".attn2.to_v.")
        names = {id(p): n for n, p in unet.named_parameters()}
        order, seen = [], set()
        self.ready_at = {}    # (kind, idx) -> number of leading flat elements complete once that block is done

        def take(module):
            for p in module.parameters():
                if id(p) in seen or not p.requires_grad or any(k in names[id(p)] for k in late_keys):
                    continue
                seen.add(id(p))
                order.append(p)

        def offset():   # every tensor starts on a multiple of 8 elements (16-byte aligned bf16 views for TMA)
            return sum((p.numel() + 7) // 8 * 8 for p in order)

        take(unet.out)
        self.ready_at[("head", 0)] = offset()
        for j in reversed(range(len(unet.output_blocks))):
            take(unet.output_blocks[j])
            self.ready_at[("out", j)] = offset()
        take(unet.middle_block)
        self.ready_at[("mid", 0)] = offset()
        for i in reversed(range(len(unet.input_blocks))):
            take(unet.input_blocks[i])
            self.ready_at[("in", i)] = offset()
        for p in unet.parameters():           # the late tensors
            if p.requires_grad and id(p) not in seen:
                seen.add(id(p))
                order.append(p)
        self.params = order
        total = offset()
        dev = order[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.offsets = []
        off = 0
        for p in order:
            self.offsets.append(off)
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += (p.numel() + 7) // 8 * 8
        step = max(1, int(bucket_mb * (1 << 20) / 4))
        self.bounds = [(lo, min(lo + step, total)) for lo in range(0, total, step)]
        self._next, self._works = 0, []
        unet.grad_ready_hook = self.block_done

    def zero_(self):
        """Use instead of optimizer.zero_grad(set_to_none=True): the views must stay attached."""
        self.flat.zero_()
        self._next, self._works = 0, []

    def _launch_upto(self, nready: int):
        while self._next < len(self.bounds) and self.bounds[self._next][1] <= nready:
            lo, hi = self.bounds[self._next]
            if self.world > 1:
                self._works.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group,
                                                   async_op=True))
            self._next += 1

    def block_done(self, kind: str, idx: int):
        self._launch_upto(self.ready_at.get((kind, idx), 0))

    def finish(self, average: bool = True):
        """Reduce whatever is left, wait for every bucket, and (unless the optimizer folds the division into its own
        pass, optim.FusedAdamWEMA.step(grads_are_sums=True)) turn the sums into means."""
        self._launch_upto(self.flat.numel())
        for w in self._works:
            w.wait()
        self._works = []
        if self.world > 1 and average:
            self.flat.mul_(1.0 / self.world)
