"""Torch-tensor front end of the C ABI: every function here is a thin argument marshaller around one
`ealdm_*` entry point of libealdm_b200.so.  PyTorch only provides device memory and the stream."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib as L


import os as _os

_NO_GN_PARTIAL = bool(_os.environ.get("EALDM_NO_GN_PARTIAL"))   # A/B switch: GroupNorm with its own statistics pass


def _dt(t: torch.dtype) -> int:
    if t == torch.float32:
        return L.F32
    if t == torch.bfloat16:
        return L.BF16
    raise TypeError(f"unsupported dtype {t}")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class Act:
    """An NHWC activation: `c` channels starting at column `c0` of a 2-D [n*h*w, ld] buffer.
    `gp` (optional) is the GroupNorm partial-statistics buffer of the WHOLE underlying buffer,
    [n*h*w/32, ld/8, 2] fp32 = {sum, sum of squares} per (32-pixel chunk, 8-channel octet): a tcgen05 convolution
    that writes this activation fills its column window of `gp`, and a GroupNorm that reads the activation then
    skips its statistics pass."""
    __slots__ = ("buf", "n", "h", "w", "c", "c0", "gp", "ln", "gunit", "lny", "gny")

    def __init__(self, buf: torch.Tensor, n: int, h: int, w: int, c: Optional[int] = None, c0: int = 0,
                 gp: Optional[torch.Tensor] = None):
        assert buf.dim() == 2 and buf.is_contiguous() and buf.shape[0] == n * h * w, \
            (tuple(buf.shape), n, h, w)
        self.buf, self.n, self.h, self.w = buf, n, h, w
        self.c0 = c0
        self.c = buf.shape[1] - c0 if c is None else c
        assert self.c0 + self.c <= buf.shape[1]
        self.gp = gp
        self.gunit = 8      # channels per entry of `gp`: octets, or quads for GroupNorm groups of 4 channels
        self.lny = None     # LayerNorm(self) written by the producing GEMM's epilogue (conv(..., ln_apply=...))
        self.gny = None     # GroupNorm(self) written by the producing conv's epilogue (conv(..., gn_apply=...))
        self.ln = None      # [rows, parts, 2] row {sum, sum of squares} partials written by the producing GEMM (ln_stats)

    @staticmethod
    def empty(n, h, w, c, dtype, device) -> "Act":
        return Act(torch.empty((n * h * w, c), dtype=dtype, device=device), n, h, w)

    @property
    def ld(self) -> int:
        return self.buf.shape[1]

    @property
    def rows(self) -> int:
        return self.n * self.h * self.w

    @property
    def ptr(self) -> int:
        return self.buf.data_ptr() + self.c0 * self.buf.element_size()

    @property
    def dtype(self) -> torch.dtype:
        return self.buf.dtype

    def cols(self, c0: int, c: int) -> "Act":
        a = Act(self.buf, self.n, self.h, self.w, c, self.c0 + c0, self.gp)
        a.gunit = self.gunit
        return a

    def images(self, i0: int, cnt: int) -> "Act":
        """Images i0 .. i0+cnt of the batch as a view (rows of the same buffer; the statistics buffer follows)."""
        assert 0 <= i0 and i0 + cnt <= self.n
        hw = self.h * self.w
        gp = None
        if self.gp is not None:
            assert hw % 32 == 0
            gp = self.gp[i0 * hw // 32:(i0 + cnt) * hw // 32]
        a = Act(self.buf[i0 * hw:(i0 + cnt) * hw], cnt, self.h, self.w, self.c, self.c0, gp)
        a.gunit = self.gunit
        return a

    def reshape(self, n, h, w) -> "Act":
        assert n * h * w == self.rows
        return Act(self.buf, n, h, w, self.c, self.c0)

    def with_gn_partial(self, unit: int = 8) -> "Act":
        """Attach a (not yet filled) partial-statistics buffer; None if the shape does not qualify.  unit = 4: one
        entry per 4 channels (GroupNorm groups of 4 channels) instead of per octet."""
        assert unit in (4, 8)
        if _NO_GN_PARTIAL:
            return self
        hw = self.h * self.w
        pow2 = lambda v: v & (v - 1) == 0  # noqa: E731
        if (hw % 32 == 0 and pow2(self.w) and pow2(self.h) and self.ld % 32 == 0 and self.c0 % 32 == 0
                and self.c % 32 == 0 and (self.w >= 32 or self.h % (32 // self.w) == 0)):
            self.gp = torch.empty((self.rows // 32, self.ld // unit, 2), dtype=torch.float32, device=self.buf.device)
            self.gunit = unit
        return self

    @property
    def gp_ptr(self) -> int:
        return self.gp.data_ptr() + (self.c0 // self.gunit) * 8

    def view2d(self) -> torch.Tensor:
        return self.buf[:, self.c0:self.c0 + self.c]


class ConvIn:
    """One source of an implicit-GEMM convolution."""
    __slots__ = ("x", "ksize", "stride", "pad", "upsample")

    def __init__(self, x: Act, ksize=1, stride=1, pad=0, upsample=0):
        self.x, self.ksize, self.stride, self.pad, self.upsample = x, ksize, stride, pad, upsample


def conv(srcs: Sequence[ConvIn], weight: torch.Tensor, out: Act, *, bias: Optional[torch.Tensor] = None,
         rowvec: Optional[torch.Tensor] = None, rowvec_col0: int = 0, residual: Optional[Act] = None,
         act: int = L.ACT_NONE, impl: int = L.IMPL_AUTO, out2: Optional[Act] = None, adjoint: bool = False,
         upsample_phases: bool = False, ln_stats: bool = False, ln: Optional[tuple] = None,
         wimg: Optional[tuple] = None, ln_apply: Optional[tuple] = None, gn_apply: Optional[tuple] = None) -> Act:
    """ealdm_conv: out = epilogue(sum_s im2col(src_s) @ weight[:, seg_s]^T). `weight` is [n_out, k_total].
    adjoint=True: data gradient of a forward layer -- `weight` is that layer's own packed matrix
    [src channels, ksize^2 * out.c] (may be a column window of a wider matrix); nothing is transposed or flipped."""
    lib = L.load()
    a = L.ConvArgs()
    x0 = srcs[0].x
    a.dtype = _dt(x0.dtype)
    a.impl = impl
    a.n_src = len(srcs)
    a.act = act
    for i, s in enumerate(srcs):
        d = a.src[i]
        d.x, d.n, d.h, d.w, d.c, d.ld = s.x.ptr, s.x.n, s.x.h, s.x.w, s.x.c, s.x.ld
        d.ksize, d.stride, d.pad, d.upsample = s.ksize, s.stride, s.pad, s.upsample
        assert s.x.dtype == x0.dtype
    a.h_out, a.w_out = out.h, out.w
    assert out.n == x0.n
    if wimg is not None:
        # per-image B operand (ealdm_conv_args::wi_*): `weight` is the context projection [n * tokens, ld] (one row per
        # context token); wimg = (first column, tokens, heads, head stride).  adjoint: out = P[., heads * tokens] @ Zt_n.
        col0, tokens, heads, hstride = wimg
        assert len(srcs) == 1 and weight.dim() == 2 and weight.dtype == x0.dtype and weight.stride(1) == 1
        assert weight.shape[0] == x0.n * tokens and col0 + heads * hstride <= weight.shape[1]
        a.weight = weight.data_ptr() + col0 * weight.element_size()
        a.wi_tokens, a.wi_heads, a.wi_ld, a.wi_head_stride = tokens, heads, weight.stride(0), hstride
        a.impl = L.IMPL_TCGEN05
        # 64-token images (8x8): two images per 128-row tile -- the logits tensor is 2 * heads * tokens wide, every row
        # holding its image's probabilities in one half and zeros in the other (ealdm_conv_args::wi_*, Params::b_img2)
        width = heads * tokens * (2 if x0.h * x0.w == 64 else 1)
        if adjoint:
            assert x0.c == width
            a.n_out, a.k_total, a.weight_adjoint = out.c, x0.c, 1
        else:
            assert out.c == width
            a.n_out, a.k_total = width, x0.c
    else:
        assert weight.dim() == 2 and weight.dtype == x0.dtype
        a.weight = weight.data_ptr()
    if wimg is not None:
        pass
    elif adjoint:
        taps = srcs[0].ksize ** 2
        assert len(srcs) == 1 and weight.stride(1) == 1 and weight.shape == (x0.c, taps * out.c), \
            (tuple(weight.shape), x0.c, taps, out.c)
        a.n_out, a.k_total = out.c, taps * x0.c
        a.weight_adjoint, a.ld_weight = 1, weight.stride(0)
    else:
        assert weight.is_contiguous()
        a.n_out, a.k_total = weight.shape
    if upsample_phases:   # `weight` = packing.pack_upsample_phases(...): four 2x2 phases over the low-resolution source
        assert len(srcs) == 1 and srcs[0].upsample == 1 and weight.shape[1] == 16 * x0.c
        a.upsample_phases = 1
        if impl == L.IMPL_AUTO:
            a.impl = L.IMPL_TCGEN05
    n_cols = a.n_out // 2 if act == L.ACT_GEGLU else a.n_out
    assert out.c == n_cols, (out.c, n_cols)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == a.n_out
        a.bias = bias.data_ptr()
    if rowvec is not None:
        assert rowvec.dtype == torch.float32 and rowvec.dim() == 2 and rowvec.shape[0] == x0.n
        a.rowvec = rowvec.data_ptr() + 4 * rowvec_col0
        a.ld_rowvec = rowvec.shape[1]
    if residual is not None:
        assert residual.rows == out.rows and residual.c == out.c
        assert residual.dtype in (x0.dtype, torch.float32)
        a.residual = residual.ptr
        a.ld_res = residual.ld
        a.res_f32 = 1 if (residual.dtype == torch.float32 and x0.dtype != torch.float32) else 0
    a.out = out.ptr
    a.ld_out = out.ld
    a.out_f32 = 1 if (out.dtype == torch.float32 and x0.dtype != torch.float32) else 0
    if not a.out_f32:
        assert out.dtype == x0.dtype
    if out2 is not None:
        assert out2.dtype == x0.dtype and out2.rows == out.rows and out2.c == out.c
        a.out2, a.ld_out2 = out2.ptr, out2.ld
    if ln_apply is not None:   # the epilogue also writes LayerNorm(out) * gamma + beta (bf16) to out2 instead of a copy
        gamma, beta, eps = ln_apply
        assert out2 is not None and a.out_f32 and act == L.ACT_NONE and out.c == 256
        assert gamma.dtype == torch.float32 and beta.dtype == torch.float32 and gamma.numel() == 256 == beta.numel()
        a.ln_gamma, a.ln_beta, a.ln_eps = gamma.data_ptr(), beta.data_ptr(), eps
        a.impl = L.IMPL_TCGEN05
    if gn_apply is not None:
        # GroupNorm (+ SiLU) applied by the epilogue: gn_apply = (gamma, beta, eps, groups, silu, only).  only=False: `out`
        # (fp32) as usual and `out2` (bf16) = act(GroupNorm(out)); only=True: `out` (bf16) = act(GroupNorm(result)) and
        # the result itself is never written.  `out.gp` (with_gn_partial) carries the statistics between the passes.
        gamma, beta, eps, groups, silu, only = gn_apply
        assert out.gp is not None and out.gunit == 8 and act == L.ACT_NONE and ln_apply is None
        assert gamma.dtype == torch.float32 and beta.dtype == torch.float32 and gamma.numel() == out.c == beta.numel()
        assert (out2 is None and not a.out_f32) if only else (out2 is not None)
        a.gn_gamma, a.gn_beta, a.gn_eps = gamma.data_ptr(), beta.data_ptr(), eps
        a.gn_groups, a.gn_silu, a.gn_only = groups, int(bool(silu)), int(bool(only))
        a.impl = L.IMPL_TCGEN05
    if out.gp is not None:     # the epilogue also produces the GroupNorm partial statistics of `out`
        assert x0.dtype == torch.bfloat16 and act != L.ACT_GEGLU
        a.gn_partial, a.gn_ld, a.gn_unit = out.gp_ptr, out.gp.shape[1], out.gunit
        if impl == L.IMPL_AUTO:
            a.impl = L.IMPL_TCGEN05   # only the tcgen05 epilogue writes them: fail loudly rather than skip silently
    if ln is not None:        # consumer of a folded LayerNorm: ln = (partials [rows, parts, 2], c1 [n_out], channels, eps)
        part, c1, channels, eps = ln
        assert part.dtype == torch.float32 and part.is_contiguous() and part.dim() == 3 and part.shape[0] == out.rows
        assert c1.dtype == torch.float32 and c1.numel() == a.n_out and bias is not None and out2 is None
        a.ln_partial_in, a.ln_parts_in, a.ln_c1 = part.data_ptr(), part.shape[1], c1.data_ptr()
        a.ln_channels, a.ln_eps = channels, eps
        a.impl = L.IMPL_TCGEN05
    if ln_stats:              # producer: the epilogue also writes every row's {sum, sum of squares} partials
        assert a.out_f32 and act == L.ACT_NONE
        a.impl = L.IMPL_TCGEN05
        a.ln_partial_out = 1      # (any non-null value: the query below only looks at the shape of the problem)
        parts = int(lib.ealdm_conv_ln_parts(C.byref(a)))
        if parts <= 0:
            L.check(-1)
        out.ln = torch.empty((out.rows, parts, 2), dtype=torch.float32, device=out.buf.device)
        a.ln_partial_out = out.ln.data_ptr()
    L.check(lib.ealdm_conv(C.byref(a), _stream()))
    return out


def linear(x: Act, weight: torch.Tensor, out: Act, **kw) -> Act:
    """nn.Linear over the rows of x (a 1x1 'convolution' with n=1, h=1, w=rows)."""
    if out.gp is not None:
        # keep the image geometry: the epilogue's GroupNorm partials are per (image, 32-pixel chunk)
        assert (x.n, x.h, x.w) == (out.n, out.h, out.w)
        return conv([ConvIn(x)], weight, out, **kw)
    xs = Act(x.buf, 1, 1, x.rows, x.c, x.c0)
    os_ = Act(out.buf, 1, 1, out.rows, out.c, out.c0)
    res = kw.pop("residual", None)
    if res is not None:
        res = Act(res.buf, 1, 1, res.rows, res.c, res.c0)
    o2 = kw.pop("out2", None)
    if o2 is not None:
        o2 = Act(o2.buf, 1, 1, o2.rows, o2.c, o2.c0)
    conv([ConvIn(xs)], weight, os_, residual=res, out2=o2, **kw)
    out.ln = os_.ln
    return out


def ff_geglu_fused(x: Act, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor, residual: Act,
                   out: Act) -> Act:
    """ealdm_ff_geglu_fused: out = GEGLU(x w1^T + b1) w2^T + b2 + residual in one kernel (c = 256, hidden = 1024);
    w1 / b1 row-interleaved by packing.geglu_interleave."""
    lib = L.load()
    a = L.FfFusedArgs()
    assert x.dtype == torch.bfloat16 and w1.dtype == torch.bfloat16 and w2.dtype == torch.bfloat16
    assert w1.is_contiguous() and w2.is_contiguous() and w1.shape == (2 * w2.shape[1], x.c) and w2.shape[0] == x.c
    assert b1.dtype == torch.float32 and b1.numel() == w1.shape[0] and b2.dtype == torch.float32 and b2.numel() == x.c
    assert residual.dtype == torch.float32 and residual.rows == x.rows and residual.c == x.c
    assert out.rows == x.rows and out.c == x.c and out.dtype in (torch.float32, torch.bfloat16)
    a.x, a.rows, a.ld_x, a.c, a.hidden = x.ptr, x.rows, x.ld, x.c, w2.shape[1]
    a.w1, a.b1, a.w2, a.b2 = w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr()
    a.residual, a.ld_res = residual.ptr, residual.ld
    a.out, a.ld_out, a.out_f32 = out.ptr, out.ld, 1 if out.dtype == torch.float32 else 0
    L.check(lib.ealdm_ff_geglu_fused(C.byref(a), _stream()))
    return out


def group_norm_workspace(n: int, hw: int, c: int, device) -> torch.Tensor:
    """Scratch for group_norm (ealdm_group_norm_workspace_bytes); reusable across calls on one stream."""
    nbytes = int(L.load().ealdm_group_norm_workspace_bytes(n, hw, c))
    return torch.empty((max(nbytes, 8) + 7) // 8, dtype=torch.float64, device=device)


def group_norm(x: Act, gamma: torch.Tensor, beta: torch.Tensor, eps: float, out: Act,
               workspace: Optional[torch.Tensor] = None, *, groups: int = 32, silu: bool = False,
               stats_out: Optional[torch.Tensor] = None) -> Act:
    lib = L.load()
    a = L.GroupNormArgs()
    a.dtype = _dt(out.dtype)
    a.x_f32 = 1 if (x.dtype == torch.float32 and out.dtype != torch.float32) else 0
    a.act = L.ACT_SILU if silu else L.ACT_NONE
    a.x, a.n, a.hw, a.c, a.ld_x = x.ptr, x.n, x.h * x.w, x.c, x.ld
    a.groups, a.eps = groups, eps
    assert gamma.dtype == torch.float32 and beta.dtype == torch.float32 and gamma.numel() == x.c
    a.gamma, a.beta = gamma.data_ptr(), beta.data_ptr()
    assert out.rows == x.rows and out.c == x.c and (out.dtype == x.dtype or x.dtype == torch.float32)
    a.y, a.ld_y = out.ptr, out.ld
    need = int(lib.ealdm_group_norm_workspace_bytes(x.n, x.h * x.w, x.c))
    if workspace is None:
        workspace = group_norm_workspace(x.n, x.h * x.w, x.c, x.buf.device)
    assert workspace.numel() * workspace.element_size() >= need, "group_norm workspace too small"
    a.workspace = workspace.data_ptr()
    if stats_out is not None:
        assert stats_out.dtype == torch.float32 and stats_out.is_contiguous() and stats_out.numel() == x.n * groups * 2
        a.stats_out = stats_out.data_ptr()
    if x.gp is not None and (x.c // groups) % x.gunit == 0:
        a.partial, a.partial_ld, a.partial_unit = x.gp_ptr, x.gp.shape[1], x.gunit
    L.check(lib.ealdm_group_norm(C.byref(a), _stream()))
    return out


def gn_partial(x: Act) -> Act:
    """Fill x.gp with a stand-alone kernel (for activations that did not come out of the tcgen05 epilogue)."""
    lib = L.load()
    assert x.gp is not None and x.gunit == 8, "the stand-alone producer writes octet partials"
    L.check(lib.ealdm_gn_partial(x.ptr, x.ld, _dt(x.dtype), x.n, x.h * x.w, x.c, x.gp_ptr, x.gp.shape[1], _stream()))
    return x


def layer_norm(x: Act, gamma: torch.Tensor, beta: torch.Tensor, eps: float, out: Act) -> Act:
    lib = L.load()
    a = L.LayerNormArgs()
    a.dtype = _dt(out.dtype)
    a.x_f32 = 1 if (x.dtype == torch.float32 and out.dtype != torch.float32) else 0
    a.x, a.rows, a.c, a.ld_x, a.eps = x.ptr, x.rows, x.c, x.ld, eps
    assert gamma.dtype == torch.float32 and gamma.numel() == x.c
    a.gamma, a.beta = gamma.data_ptr(), beta.data_ptr()
    assert out.rows == x.rows and out.c == x.c and (out.dtype == x.dtype or x.dtype == torch.float32)
    a.y, a.ld_y = out.ptr, out.ld
    L.check(lib.ealdm_layer_norm(C.byref(a), _stream()))
    return out


def attention(q: Act, k: Act, v: Act, out: Act, *, batch: int, heads: int, head_dim: int, n_q: int,
              n_kv: int, scale: float, head_stride_q: Optional[int] = None,
              head_stride_kv: Optional[int] = None, impl: int = L.IMPL_AUTO,
              lse: Optional[torch.Tensor] = None) -> Act:
    """q/k/v are column windows into [batch*n, ld] buffers; head h of q starts at column h*head_stride_q."""
    lib = L.load()
    a = L.AttentionArgs()
    a.dtype = _dt(q.dtype)
    a.impl = impl
    a.q, a.k, a.v = q.ptr, k.ptr, v.ptr
    assert k.ld == v.ld and q.rows == batch * n_q and k.rows == batch * n_kv
    a.ld_q, a.ld_kv = q.ld, k.ld
    a.head_stride_q = head_dim if head_stride_q is None else head_stride_q
    a.head_stride_kv = head_dim if head_stride_kv is None else head_stride_kv
    a.batch, a.heads, a.n_q, a.n_kv, a.head_dim = batch, heads, n_q, n_kv, head_dim
    a.scale = scale
    assert out.rows == q.rows and out.c == heads * head_dim and out.dtype == q.dtype
    a.out, a.ld_out = out.ptr, out.ld
    if lse is not None:
        assert lse.dtype == torch.float32 and lse.is_contiguous() and lse.numel() == batch * heads * n_q
        a.lse = lse.data_ptr()
    L.check(lib.ealdm_attention(C.byref(a), _stream()))
    return out


def timestep_embedding(t: torch.Tensor, dim: int, freqs: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    lib = L.load()
    assert t.dtype == torch.int64 and t.is_contiguous() and freqs.dtype == torch.float32
    assert out.is_contiguous() and tuple(out.shape) == (t.numel(), dim)
    L.check(lib.ealdm_timestep_embedding(t.data_ptr(), t.numel(), dim, freqs.data_ptr(), _dt(out.dtype),
                                         out.data_ptr(), _stream()))
    return out


def nchw_to_nhwc(x: torch.Tensor, out: Act) -> Act:
    lib = L.load()
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 4
    n, c, h, w = x.shape
    assert (out.n, out.h, out.w, out.c) == (n, h, w, c)
    L.check(lib.ealdm_nchw_to_nhwc(x.data_ptr(), n, c, h, w, _dt(out.dtype), out.ptr, out.ld, _stream()))
    return out


def nhwc_to_nchw(x: Act, out: torch.Tensor) -> torch.Tensor:
    lib = L.load()
    assert out.dtype == torch.float32 and out.is_contiguous()
    assert tuple(out.shape) == (x.n, x.c, x.h, x.w)
    L.check(lib.ealdm_nhwc_to_nchw(x.ptr, x.ld, _dt(x.dtype), x.n, x.c, x.h, x.w, out.data_ptr(), _stream()))
    return out


def upsample_nearest2x(x: Act, out: Act) -> Act:
    lib = L.load()
    assert (out.n, out.h, out.w, out.c) == (x.n, 2 * x.h, 2 * x.w, x.c) and out.dtype == x.dtype
    L.check(lib.ealdm_upsample_nearest2x(x.ptr, x.ld, _dt(x.dtype), x.n, x.h, x.w, x.c, out.ptr, out.ld,
                                         _stream()))
    return out


def copy2d(x: Act, out: Act) -> Act:
    lib = L.load()
    assert x.rows == out.rows and x.c == out.c
    L.check(lib.ealdm_copy2d(x.ptr, x.ld, _dt(x.dtype), out.ptr, out.ld, _dt(out.dtype), x.rows, x.c,
                             _stream()))
    return out


def softmax_rows_(x: Act, scale: float) -> Act:
    lib = L.load()
    L.check(lib.ealdm_softmax_rows(x.ptr, x.ld, _dt(x.dtype), x.rows, x.c, scale, _stream()))
    return x


def ddim_step(x, e_cond, *, e_uncond=None, noise=None, cfg_scale=1.0, sqrt_one_minus_at, sqrt_at,
              sqrt_a_prev, dir_coef, sigma_t, temperature=1.0, want_e=False):
    """ealdm_ddim_step on contiguous fp32 tensors; returns (x_prev, pred_x0[, e])."""
    lib = L.load()
    for t in (x, e_cond, e_uncond, noise):
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous() and t.numel() == x.numel())
    x_prev = torch.empty_like(x)
    pred = torch.empty_like(x)
    e = torch.empty_like(x) if want_e else None
    a = L.DdimStepArgs()
    a.x, a.e_uncond, a.e_cond, a.noise = _ptr(x), _ptr(e_uncond), _ptr(e_cond), _ptr(noise)
    a.x_prev, a.pred_x0, a.e_out = _ptr(x_prev), _ptr(pred), _ptr(e)
    a.numel = x.numel()
    a.cfg_scale = cfg_scale
    a.sqrt_one_minus_at, a.sqrt_at, a.sqrt_a_prev = sqrt_one_minus_at, sqrt_at, sqrt_a_prev
    a.dir_coef, a.sigma_t, a.temperature = dir_coef, sigma_t, temperature
    L.check(lib.ealdm_ddim_step(C.byref(a), _stream()))
    return (x_prev, pred, e) if want_e else (x_prev, pred)


def plms_eps(e_cond, *, e_uncond=None, cfg_scale=1.0, old=(), mode=0):
    """ealdm_plms_eps: returns (e_t, e_prime); `old` = (old_eps[-1], old_eps[-2], old_eps[-3]) as far as needed."""
    lib = L.load()
    for t in (e_cond, e_uncond) + tuple(old):
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous() and t.numel() == e_cond.numel())
    e_t, e_p = torch.empty_like(e_cond), torch.empty_like(e_cond)
    o = list(old) + [None] * (3 - len(old))
    L.check(lib.ealdm_plms_eps(_ptr(e_uncond), e_cond.data_ptr(), cfg_scale, _ptr(o[0]), _ptr(o[1]), _ptr(o[2]), mode,
                               e_t.data_ptr(), e_p.data_ptr(), e_cond.numel(), _stream()))
    return e_t, e_p


def ddpm_step(x, eps, noise, t, bufs, *, clip_denoised=False, temperature=1.0, want_x0=False):
    """ealdm_ddpm_step; `bufs` = (sqrt_recip_alphas_cumprod, sqrt_recipm1_alphas_cumprod, posterior_mean_coef1,
    posterior_mean_coef2, posterior_log_variance_clipped), fp32 device tensors."""
    lib = L.load()
    for v in (x, eps, noise) + tuple(bufs):
        assert v.dtype == torch.float32 and v.is_contiguous() and v.is_cuda
    assert t.dtype == torch.int64 and t.is_contiguous()
    x_prev = torch.empty_like(x)
    x0 = torch.empty_like(x) if want_x0 else None
    b = x.shape[0]
    L.check(lib.ealdm_ddpm_step(x.data_ptr(), eps.data_ptr(), noise.data_ptr(), t.data_ptr(), *[v.data_ptr() for v in bufs],
                                1 if clip_denoised else 0, float(temperature), b, x.numel() // b, x_prev.data_ptr(),
                                _ptr(x0), _stream()))
    return (x_prev, x0) if want_x0 else x_prev


def q_sample(x0, noise, t, sqrt_ac, sqrt_1mac):
    lib = L.load()
    assert x0.dtype == torch.float32 and x0.is_contiguous() and noise.is_contiguous()
    assert t.dtype == torch.int64 and sqrt_ac.dtype == torch.float32 and sqrt_1mac.dtype == torch.float32
    out = torch.empty_like(x0)
    b = x0.shape[0]
    L.check(lib.ealdm_q_sample(x0.data_ptr(), noise.data_ptr(), t.data_ptr(), sqrt_ac.data_ptr(),
                               sqrt_1mac.data_ptr(), b, x0.numel() // b, out.data_ptr(), _stream()))
    return out


def cfg_mse(e_cond, target, *, e_uncond=None, cfg_scale=1.0):
    lib = L.load()
    for t in (e_cond, target, e_uncond):
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous())
    b = e_cond.shape[0]
    out = torch.empty((b,), dtype=torch.float32, device=e_cond.device)
    L.check(lib.ealdm_cfg_mse(_ptr(e_uncond), e_cond.data_ptr(), target.data_ptr(), cfg_scale, b,
                              e_cond.numel() // b, out.data_ptr(), _stream()))
    return out


# ---- backward pass ------------------------------------------------------------------------------------
class Workspace:
    """A grow-only scratch buffer for the backward kernels (one per stream of execution)."""

    def __init__(self, device):
        self.device, self.buf = device, None

    def get(self, nbytes: int) -> torch.Tensor:
        n = (max(int(nbytes), 16) + 15) // 16 * 2
        if self.buf is None or self.buf.numel() < n:
            self.buf = torch.empty(n, dtype=torch.float64, device=self.device)
        return self.buf


def conv_wgrad(x: Act, dy: Act, dw: torch.Tensor, ws: Workspace, *, ksize=1, stride=1, pad=0, col0: int = 0,
               layout: int = L.WGRAD_OIHW, accumulate: bool = True, impl: int = L.IMPL_AUTO) -> None:
    """ealdm_conv_wgrad: dw (+)= dy^T * im2col(x).  `dw` is a 2-D fp32 view [n_out, ld]; with the PACKED layout
    the result goes to columns [col0, col0 + ksize^2 * c)."""
    lib = L.load()
    a = L.ConvWgradArgs()
    a.dtype, a.impl = _dt(x.dtype), impl
    d = a.src
    d.x, d.n, d.h, d.w, d.c, d.ld = x.ptr, x.n, x.h, x.w, x.c, x.ld
    d.ksize, d.stride, d.pad, d.upsample = ksize, stride, pad, 0
    assert dy.dtype == x.dtype and dy.n == x.n
    a.dy, a.ld_dy, a.n_out, a.h_out, a.w_out = dy.ptr, dy.ld, dy.c, dy.h, dy.w
    assert dw.dtype == torch.float32 and dw.dim() == 2 and dw.stride(1) == 1 and dw.shape[0] == dy.c
    a.dw = dw.data_ptr() + 4 * col0
    a.ld_dw = dw.stride(0)
    assert col0 + ksize * ksize * x.c <= dw.shape[1]
    a.layout, a.accumulate = layout, 1 if accumulate else 0
    need = int(lib.ealdm_conv_wgrad_workspace_bytes(C.byref(a)))
    if need < 0:
        L.check(-1)
    w = ws.get(need)
    a.workspace, a.workspace_bytes = w.data_ptr(), w.numel() * 8
    L.check(lib.ealdm_conv_wgrad(C.byref(a), _stream()))


def linear_wgrad(x: Act, dy: Act, dw: torch.Tensor, ws: Workspace, **kw) -> None:
    xs = Act(x.buf, 1, 1, x.rows, x.c, x.c0)
    ds = Act(dy.buf, 1, 1, dy.rows, dy.c, dy.c0)
    conv_wgrad(xs, ds, dw, ws, **kw)


def group_norm_bwd(x: Act, dy: Act, stats: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, dx: Act,
                   ws: Workspace, *, groups: int = 32, silu: bool = False, add: Optional[Act] = None,
                   add2: Optional[Act] = None, dx2: Optional[Act] = None, dgamma: Optional[torch.Tensor] = None,
                   dbeta: Optional[torch.Tensor] = None) -> Act:
    lib = L.load()
    a = L.GroupNormBwdArgs()
    a.dtype = _dt(dy.dtype)
    a.act = L.ACT_SILU if silu else L.ACT_NONE
    a.x, a.n, a.hw, a.c, a.ld_x = x.ptr, x.n, x.h * x.w, x.c, x.ld
    a.groups = groups
    a.x_f32 = 1 if (x.dtype == torch.float32 and dy.dtype != torch.float32) else 0
    assert stats.dtype == torch.float32 and stats.numel() == x.n * groups * 2
    a.stats, a.gamma, a.beta = stats.data_ptr(), gamma.data_ptr(), beta.data_ptr()
    assert dy.rows == x.rows and dy.c == x.c and dx.rows == x.rows and dx.c == x.c
    a.dy, a.ld_dy = dy.ptr, dy.ld
    for name, t in (("add", add), ("add2", add2)):
        if t is not None:
            assert t.dtype == torch.float32 and t.rows == x.rows and t.c == x.c
            setattr(a, name, t.ptr)
            setattr(a, "ld_" + name, t.ld)
    a.dx, a.ld_dx = dx.ptr, dx.ld
    a.dx_f32 = 1 if (dx.dtype == torch.float32 and dy.dtype != torch.float32) else 0
    assert dx.dtype in (dy.dtype, torch.float32)
    if dx2 is not None:
        assert dx2.dtype == dy.dtype and dx2.rows == x.rows and dx2.c == x.c
        a.dx2, a.ld_dx2 = dx2.ptr, dx2.ld
    for name, t in (("dgamma", dgamma), ("dbeta", dbeta)):
        if t is not None:
            assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == x.c
            setattr(a, name, t.data_ptr())
    a.workspace = ws.get(int(lib.ealdm_group_norm_bwd_workspace_bytes(x.n, x.h * x.w, x.c))).data_ptr()
    L.check(lib.ealdm_group_norm_bwd(C.byref(a), _stream()))
    return dx


def layer_norm_bwd(x: Act, dy: Act, gamma: torch.Tensor, eps: float, dx: Act, ws: Workspace, *,
                   add: Optional[Act] = None, dx2: Optional[Act] = None, dgamma: Optional[torch.Tensor] = None,
                   dbeta: Optional[torch.Tensor] = None) -> Act:
    lib = L.load()
    a = L.LayerNormBwdArgs()
    a.dtype = _dt(dy.dtype)
    a.x_f32 = 1 if (x.dtype == torch.float32 and dy.dtype != torch.float32) else 0
    a.dx_f32 = 1 if (dx.dtype == torch.float32 and dy.dtype != torch.float32) else 0
    a.x, a.rows, a.c, a.ld_x, a.eps = x.ptr, x.rows, x.c, x.ld, eps
    a.gamma = gamma.data_ptr()
    assert dy.rows == x.rows and dy.c == x.c and dx.rows == x.rows and dx.c == x.c
    a.dy, a.ld_dy, a.dx, a.ld_dx = dy.ptr, dy.ld, dx.ptr, dx.ld
    if add is not None:
        assert add.dtype == torch.float32 and add.rows == x.rows and add.c == x.c
        a.add, a.ld_add = add.ptr, add.ld
    if dx2 is not None:
        assert dx2.dtype == dy.dtype
        a.dx2, a.ld_dx2 = dx2.ptr, dx2.ld
    for name, t in (("dgamma", dgamma), ("dbeta", dbeta)):
        if t is not None:
            assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == x.c
            setattr(a, name, t.data_ptr())
    a.workspace = ws.get(int(lib.ealdm_layer_norm_bwd_workspace_bytes(x.rows, x.c))).data_ptr()
    L.check(lib.ealdm_layer_norm_bwd(C.byref(a), _stream()))
    return dx


def attention_bwd(q: Act, k: Act, v: Act, out: Act, dout: Act, dq: Act, dk: Act, dv: Act, ws: Workspace, *,
                  batch: int, heads: int, head_dim: int, n_q: int, n_kv: int, scale: float,
                  head_stride_q: Optional[int] = None, head_stride_kv: Optional[int] = None,
                  head_stride_dq: Optional[int] = None, head_stride_dkv: Optional[int] = None,
                  lse: Optional[torch.Tensor] = None, impl: int = L.IMPL_AUTO) -> None:
    lib = L.load()
    a = L.AttentionBwdArgs()
    a.dtype = _dt(q.dtype)
    a.q, a.k, a.v, a.out, a.dout = q.ptr, k.ptr, v.ptr, out.ptr, dout.ptr
    assert k.ld == v.ld and dk.ld == dv.ld and q.rows == batch * n_q and k.rows == batch * n_kv
    a.ld_q, a.ld_kv, a.ld_out, a.ld_dout = q.ld, k.ld, out.ld, dout.ld
    a.head_stride_q = head_dim if head_stride_q is None else head_stride_q
    a.head_stride_kv = head_dim if head_stride_kv is None else head_stride_kv
    a.batch, a.heads, a.n_q, a.n_kv, a.head_dim, a.scale = batch, heads, n_q, n_kv, head_dim, scale
    a.dq, a.ld_dq = dq.ptr, dq.ld
    a.head_stride_dq = head_dim if head_stride_dq is None else head_stride_dq
    a.dk, a.dv, a.ld_dkv = dk.ptr, dv.ptr, dk.ld
    a.head_stride_dkv = head_dim if head_stride_dkv is None else head_stride_dkv
    a.impl = impl
    if lse is not None:
        assert lse.dtype == torch.float32 and lse.is_contiguous() and lse.numel() == batch * heads * n_q
        a.lse = lse.data_ptr()
    need = int(lib.ealdm_attention_bwd_workspace_bytes(C.byref(a)))
    w = ws.get(need)
    a.workspace, a.workspace_bytes = w.data_ptr(), w.numel() * 8
    L.check(lib.ealdm_attention_bwd(C.byref(a), _stream()))


def geglu(pre: Act, out: Act) -> Act:
    lib = L.load()
    assert pre.c == 2 * out.c and pre.rows == out.rows and pre.dtype == out.dtype
    L.check(lib.ealdm_geglu(pre.ptr, pre.ld, _dt(pre.dtype), pre.rows, out.c, out.ptr, out.ld, _stream()))
    return out


def geglu_bwd(pre: Act, dout: Act, dpre: Act) -> Act:
    lib = L.load()
    assert pre.c == 2 * dout.c == dpre.c and pre.dtype == dout.dtype == dpre.dtype
    L.check(lib.ealdm_geglu_bwd(pre.ptr, pre.ld, dout.ptr, dout.ld, _dt(pre.dtype), pre.rows, dout.c, dpre.ptr,
                                dpre.ld, _stream()))
    return dpre


def silu(x: Act, out: Act) -> Act:
    lib = L.load()
    assert x.dtype == torch.float32 and x.rows == out.rows and x.c == out.c
    L.check(lib.ealdm_silu(x.ptr, x.ld, x.rows, x.c, _dt(out.dtype), out.ptr, out.ld, _stream()))
    return out


def silu_bwd(x: Act, dy: Act, dx: Act) -> Act:
    lib = L.load()
    assert x.dtype == torch.float32 and dy.dtype == dx.dtype and x.rows == dy.rows == dx.rows
    L.check(lib.ealdm_silu_bwd(x.ptr, x.ld, dy.ptr, dy.ld, _dt(dy.dtype), x.rows, x.c, dx.ptr, dx.ld, _stream()))
    return dx


def colsum(x: Act, out: torch.Tensor, ws: Workspace, *, segs: int = 1, col0: int = 0,
           accumulate: bool = True) -> torch.Tensor:
    """out[s, col0:col0+c] (+)= sum over the rows of segment s of x (segs equal row segments)."""
    lib = L.load()
    assert out.dtype == torch.float32 and x.rows % segs == 0
    o2 = out if out.dim() == 2 else out.view(1, -1)
    assert o2.shape[0] == segs and o2.stride(1) == 1 and col0 + x.c <= o2.shape[1]
    w = ws.get(int(lib.ealdm_colsum_workspace_bytes(segs, x.rows // segs, x.c)))
    L.check(lib.ealdm_colsum(x.ptr, x.ld, _dt(x.dtype), segs, x.rows // segs, x.c, o2.data_ptr() + 4 * col0,
                             o2.stride(0), 1 if accumulate else 0, w.data_ptr(), _stream()))
    return out


def zero_insert2x(dy: Act, z: Act) -> Act:
    lib = L.load()
    assert (z.n, z.h, z.w, z.c) == (dy.n, 2 * dy.h, 2 * dy.w, dy.c) and z.dtype == dy.dtype
    L.check(lib.ealdm_zero_insert2x(dy.ptr, dy.ld, _dt(dy.dtype), dy.n, dy.h, dy.w, dy.c, z.ptr, z.ld, _stream()))
    return z


def sumpool2x2(dup: Act, dx: Act, *, add: Optional[Act] = None, dx2: Optional[Act] = None) -> Act:
    lib = L.load()
    assert (dup.n, dup.h, dup.w, dup.c) == (dx.n, 2 * dx.h, 2 * dx.w, dx.c) and dx.dtype == torch.float32
    if add is not None:
        assert add.dtype == torch.float32 and add.rows == dx.rows and add.c == dx.c
    if dx2 is not None:
        assert dx2.dtype == dup.dtype
    L.check(lib.ealdm_sumpool2x2(dup.ptr, dup.ld, _dt(dup.dtype), dx.n, dx.h, dx.w, dx.c,
                                 None if add is None else add.ptr, 0 if add is None else add.ld, dx.ptr, dx.ld,
                                 None if dx2 is None else dx2.ptr, 0 if dx2 is None else dx2.ld, _stream()))
    return dx


def cfg_mse_bwd(e_cond, target, w, *, e_uncond=None, cfg_scale=1.0):
    """Gradients of sum_b w[b] * loss_simple[b] w.r.t. (e_uncond, e_cond)."""
    lib = L.load()
    for t in (e_cond, target, e_uncond, w):
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous())
    b = e_cond.shape[0]
    de_c = torch.empty_like(e_cond)
    de_u = torch.empty_like(e_cond) if e_uncond is not None else None
    L.check(lib.ealdm_cfg_mse_bwd(_ptr(e_uncond), e_cond.data_ptr(), target.data_ptr(), w.data_ptr(), cfg_scale, b,
                                  e_cond.numel() // b, _ptr(de_u), de_c.data_ptr(), _stream()))
    return de_u, de_c


def vq_nearest(z: torch.Tensor, codebook: torch.Tensor):
    """ealdm_vq_nearest on an NCHW fp32 latent; returns (z_q NCHW fp32, indices int64 [n*h*w])."""
    lib = L.load()
    assert z.dtype == torch.float32 and z.is_contiguous() and z.dim() == 4
    assert codebook.dtype == torch.float32 and codebook.is_contiguous() and codebook.shape[1] == z.shape[1]
    n, e, h, w = z.shape
    zq = torch.empty_like(z)
    idx = torch.empty((n * h * w,), dtype=torch.int64, device=z.device)
    L.check(lib.ealdm_vq_nearest(z.data_ptr(), n, e, h * w, codebook.data_ptr(), codebook.shape[0], zq.data_ptr(),
                                 idx.data_ptr(), _stream()))
    return zq, idx


def launch_count() -> int:
    return int(L.load().ealdm_launch_count())


# ---- EALDM conditioner (csrc/cond.cu) ------------------------------------------------------------------------
def fourier_style(time: torch.Tensor, freqs: torch.Tensor, include_lin: bool, lin_lr: float, weight: torch.Tensor,
                  gain: float, out: torch.Tensor, features: Optional[torch.Tensor] = None) -> torch.Tensor:
    """ealdm_fourier_style: ConditioningTransform + CondScale of the time stamp -> out [T, n_out] (fp32)."""
    lib = L.load()
    for t in (time, freqs, weight, out):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.is_cuda
    T, nf = time.numel(), freqs.numel()
    assert weight.shape[1] == 2 * nf and tuple(out.shape) == (T, weight.shape[0])
    if features is not None:
        assert features.dtype == torch.float32 and features.is_contiguous() and tuple(features.shape) == (T, 2 * nf)
    L.check(lib.ealdm_fourier_style(time.data_ptr(), T, freqs.data_ptr(), nf, int(include_lin), float(lin_lr),
                                    weight.data_ptr(), weight.shape[0], float(gain), _ptr(features), out.data_ptr(),
                                    _stream()))
    return out


def lstm_cell(x: torch.Tensor, w_ih: torch.Tensor, b_ih: torch.Tensor, b_hh: torch.Tensor, h_out: torch.Tensor,
              c_out: torch.Tensor, rec: Optional[torch.Tensor] = None, c_prev: Optional[torch.Tensor] = None) -> None:
    """ealdm_lstm_cell: one nn.LSTM step for x [B, in] (a strided view is fine); rec = W_hh h_prev [B, 4H] or None."""
    lib = L.load()
    B, n_in = x.shape
    H = w_ih.shape[0] // 4
    assert x.dtype == torch.float32 and x.stride(1) == 1 and w_ih.is_contiguous() and w_ih.shape[1] == n_in
    assert h_out.stride(1) == 1 and tuple(h_out.shape) == (B, H) and c_out.is_contiguous() and tuple(c_out.shape) == (B, H)
    if rec is not None:
        assert rec.is_contiguous() and tuple(rec.shape) == (B, 4 * H)
    if c_prev is not None:
        assert c_prev.is_contiguous() and tuple(c_prev.shape) == (B, H)
    L.check(lib.ealdm_lstm_cell(x.data_ptr(), x.stride(0), B, n_in, w_ih.data_ptr(), b_ih.data_ptr(), b_hh.data_ptr(),
                                _ptr(rec), _ptr(c_prev), H, h_out.data_ptr(), h_out.stride(0), c_out.data_ptr(),
                                _stream()))


def adain(x: Act, style: torch.Tensor, out: Act, eps: float = 1e-5) -> Act:
    """ealdm_adain: instance norm + x * (1 + gamma) + beta, style [n, 2c] = [gamma | beta]; fp32, c <= 32."""
    lib = L.load()
    assert x.dtype == torch.float32 and out.dtype == torch.float32 and x.c == out.c and x.rows == out.rows
    assert style.dtype == torch.float32 and style.stride(1) == 1 and tuple(style.shape) == (x.n, 2 * x.c)
    L.check(lib.ealdm_adain(x.ptr, x.ld, x.n, x.h * x.w, x.c, style.data_ptr(), style.stride(0), float(eps), out.ptr,
                            out.ld, _stream()))
    return out


def batch_norm_relu(x: Act, gamma: torch.Tensor, beta: torch.Tensor, running_mean: torch.Tensor,
                    running_var: torch.Tensor, out: Act, *, training: bool, eps: float = 1e-5, relu: bool = True,
                    batch_stats: Optional[torch.Tensor] = None) -> Act:
    """ealdm_batch_norm_relu over the rows of an NHWC map (c <= 32); training=True: statistics of the batch, written
    to batch_stats [2, c] = {mean, biased variance} when given."""
    lib = L.load()
    assert x.dtype == torch.float32 and out.dtype == torch.float32 and x.c == out.c and x.rows == out.rows
    if batch_stats is not None:
        assert batch_stats.dtype == torch.float32 and batch_stats.is_contiguous() and batch_stats.numel() == 2 * x.c
    L.check(lib.ealdm_batch_norm_relu(x.ptr, x.ld, x.rows, x.c, gamma.data_ptr(), beta.data_ptr(),
                                      running_mean.data_ptr(), running_var.data_ptr(), int(training), float(eps),
                                      int(relu), out.ptr, out.ld, _ptr(batch_stats), _stream()))
    return out


def adain_bwd(x: Act, dy: Act, dstyle: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """ealdm_adain_bwd: style gradient [n, 2c] = [dgamma | dbeta] of AdaIN (its input x is frozen)."""
    lib = L.load()
    assert x.dtype == torch.float32 and dy.dtype == torch.float32 and x.c == dy.c and x.rows == dy.rows
    assert dstyle.dtype == torch.float32 and dstyle.stride(1) == 1 and tuple(dstyle.shape) == (x.n, 2 * x.c)
    L.check(lib.ealdm_adain_bwd(x.ptr, x.ld, x.n, x.h * x.w, x.c, dy.ptr, dy.ld, float(eps), dstyle.data_ptr(),
                                dstyle.stride(0), _stream()))
    return dstyle


def batch_norm_relu_bwd(x: Act, y: Act, dy: Act, dx: Act, gamma: torch.Tensor, running_mean: torch.Tensor,
                        running_var: torch.Tensor, *, training: bool, eps: float = 1e-5, relu: bool = True,
                        dgamma: Optional[torch.Tensor] = None, dbeta: Optional[torch.Tensor] = None) -> Act:
    """ealdm_batch_norm_relu_bwd; dgamma / dbeta are accumulated into."""
    lib = L.load()
    for t in (x, y, dy, dx):
        assert t.dtype == torch.float32 and t.rows == x.rows and t.c == x.c
    L.check(lib.ealdm_batch_norm_relu_bwd(x.ptr, x.ld, x.rows, x.c, gamma.data_ptr(), running_mean.data_ptr(),
                                          running_var.data_ptr(), int(training), float(eps), int(relu), y.ptr, y.ld,
                                          dy.ptr, dy.ld, dx.ptr, dx.ld, _ptr(dgamma), _ptr(dbeta), _stream()))
    return dx


def lstm_cell_bwd(x: torch.Tensor, w_ih: torch.Tensor, b_ih: torch.Tensor, b_hh: torch.Tensor, dh: torch.Tensor,
                  dgates: torch.Tensor, rec: Optional[torch.Tensor] = None, c_prev: Optional[torch.Tensor] = None,
                  dc_next: Optional[torch.Tensor] = None, dc_prev: Optional[torch.Tensor] = None) -> torch.Tensor:
    """ealdm_lstm_cell_bwd: dgates [B, 4H] (pre-activation) of one LSTM step; dc_prev [B, H] when given."""
    lib = L.load()
    B, n_in = x.shape
    H = w_ih.shape[0] // 4
    assert x.dtype == torch.float32 and x.stride(1) == 1 and dh.stride(1) == 1 and tuple(dh.shape) == (B, H)
    assert dgates.is_contiguous() and tuple(dgates.shape) == (B, 4 * H)
    for t in (rec, c_prev, dc_next, dc_prev):
        assert t is None or t.is_contiguous()
    L.check(lib.ealdm_lstm_cell_bwd(x.data_ptr(), x.stride(0), B, n_in, w_ih.data_ptr(), b_ih.data_ptr(), b_hh.data_ptr(),
                                    _ptr(rec), _ptr(c_prev), H, dh.data_ptr(), dh.stride(0), _ptr(dc_next),
                                    dgates.data_ptr(), _ptr(dc_prev), _stream()))
    return dgates


def relu_bwd(y: torch.Tensor, dy: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """ealdm_relu_bwd on 2-D fp32 tensors: dy * (y > 0) [* mask]."""
    lib = L.load()
    assert y.dim() == 2 and y.dtype == torch.float32 and y.stride(1) == 1 and dy.shape == y.shape and dy.stride(1) == 1
    dx = torch.empty_like(y)
    if mask is not None:
        assert mask.shape == y.shape and mask.dtype == torch.float32 and mask.stride(1) == 1
    L.check(lib.ealdm_relu_bwd(y.data_ptr(), y.stride(0), dy.data_ptr(), dy.stride(0), _ptr(mask),
                               0 if mask is None else mask.stride(0), y.shape[0], y.shape[1], dx.data_ptr(),
                               dx.stride(0), _stream()))
    return dx
