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


def pack_upsample_phases(w: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """`Upsample` (openaimodel.py:109-119) = nearest-2x then conv3x3 pad 1.  Output pixel (2y + py, 2x + px) reads
    upsampled rows 2y + py + {-1, 0, 1}, i.e. SOURCE rows {y - 1, y, y} for py = 0 and {y, y, y + 1} for py = 1 (same
    for columns): every output phase is a 2x2 convolution over the source whose taps are sums of the 3x3 taps that
    land on the same source pixel.  [O, I, 3, 3] -> [O, (py, px, row tap, column tap, I)] = [O, 16 I]; the sums are
    formed in fp32 and rounded once to `dtype`.  Tap (a, b) of phase (py, px) reads source pixel (y + py - 1 + a,
    x + px - 1 + b); out-of-range source pixels are the zero padding of the upsampled image."""
    assert w.dim() == 4 and w.shape[2:] == (3, 3)
    wf = w.detach().float()
    # rows of the 3x3 kernel (dy = -1, 0, +1) that fall on source-row tap a, per phase
    groups = {0: ((0,), (1, 2)), 1: ((0, 1), (2,))}
    phases = []
    for py in (0, 1):
        for px in (0, 1):
            taps = []
            for a in (0, 1):
                for b in (0, 1):
                    acc = torch.zeros_like(wf[:, :, 0, 0])
                    for r in groups[py][a]:
                        for c in groups[px][b]:
                            acc = acc + wf[:, :, r, c]
                    taps.append(acc)                      # [O, I]
            phases.append(torch.stack(taps, dim=1))        # [O, 4, I]
    return torch.stack(phases, dim=1).reshape(w.shape[0], -1).to(dtype).contiguous()   # [O, 4 * 4 * I]
