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


def collapse_cross_attention(wq: torch.Tensor, wk: torch.Tensor, wv: torch.Tensor, wo: torch.Tensor, heads: int,
                             scale: float, dtype: torch.dtype):
    """CrossAttention (attention.py:170-193) on a context of a few tokens, collapsed onto the context.  With
    q = x Wq^T, k_j = ctx_j Wk^T, v_j = ctx_j Wv^T the logits and the output projection are bilinear in (x, ctx):
        sim[m, h, j]  = scale * sum_{d in head h} q[m, d] k_j[d]  =  x[m, :] . (ctx_j G_h),   G_h = scale Wk_h^T Wq_h
        to_out(attn @ v)[m, :] = sum_{h, j} p[m, h, j] (ctx_j H_h) + b,                       H_h = Wv_h^T Wo_h^T
    so ONE projection of the context by [G; H] (weights only, formed here in fp32 and rounded once) yields, per image,
    the [heads * tokens, C] matrices the two per-image GEMMs of ealdm_conv(wi_*) multiply by.  log2(e) is folded into G:
    the kernel's softmax runs in base 2.  Returns (G, H), each [heads * C, context_dim] with rows ordered (head, c)."""
    inner, c = wq.shape
    d = inner // heads
    ctx_dim = wk.shape[1]
    wq3 = wq.detach().float().reshape(heads, d, c)
    wk3 = wk.detach().float().reshape(heads, d, ctx_dim)
    wv3 = wv.detach().float().reshape(heads, d, ctx_dim)
    wo3 = wo.detach().float().reshape(wo.shape[0], heads, d)
    g = torch.einsum("hdc,hde->hce", wq3, wk3) * (scale * 1.4426950408889634)
    h = torch.einsum("chd,hde->hce", wo3, wv3)
    return (g.reshape(heads * c, ctx_dim).to(dtype).contiguous(),
            h.reshape(heads * wo.shape[0], ctx_dim).to(dtype).contiguous())
