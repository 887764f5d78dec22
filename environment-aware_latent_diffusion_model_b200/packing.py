"""Weight repacking from the reference's state_dict layout to the kernel layouts.

The canonical parameters stay fp32 `nn.Parameter`s with the reference's names and shapes (so
reference checkpoints load); the packed copies are derived lazily and cached by the modules."""
from __future__ import annotations

import torch


def pack_conv_weight(w: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """[O, I, kh, kw] (nn.Conv2d) -> [O, kh*kw*I] with K ordered (kh, kw, c), matching NHWC gathers."""
    if w.dim() == 3:  # conv1d 1x1 (AttentionBlock.qkv / proj_out, openaimodel.py:304,312)
        w = w[..., None]
    o = w.shape[0]
    # cast first: the permuting copy then moves half the bytes
    return w.detach().to(dtype).permute(0, 2, 3, 1).reshape(o, -1).contiguous()


def geglu_interleave(w: torch.Tensor, b: torch.Tensor):
    """GEGLU.proj (attention.py:38-44) computes [value | gate] = x W^T + b and returns
    value * gelu(gate).  Rows are permuted so that every block of 32 accumulator columns holds 16
    value columns followed by their 16 gate columns: the kernel epilogue then owns complete pairs."""
    n2 = w.shape[0]
    half = n2 // 2
    assert half % 16 == 0
    p = torch.arange(n2, device=w.device)
    blk, within = p // 32, p % 32
    j = blk * 16 + within % 16
    src = torch.where(within < 16, j, half + j)
    return w.detach()[src].contiguous(), b.detach()[src].float().contiguous()
