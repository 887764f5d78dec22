"""Training-mode execution of the UNet: forward with saved activations and the hand-written backward
(reference: `loss.backward()` through ldm/modules/diffusionmodules/openaimodel.py and ldm/modules/attention.py
inside LatentDiffusion.p_losses, ldm/models/diffusion/ddpm.py:1036-1078).

The reference recomputes every transformer block in the backward (`checkpoint`, attention.py:209,
util.py:119-148) because its N^2 attention scores do not fit; here nothing N^2 is ever stored, so all
activations are simply kept (a few GB of the 180 GB at batch 2x32) and nothing is recomputed.

`UNetTrainEngine.forward` runs the same launch sequence as `UNetEngine.forward` except that
  * GroupNorm also writes its (mean, rstd), and every operand needed by a gradient GEMM is kept;
  * the GEGLU feed-forward keeps its pre-activation (natural [value | gate] layout) instead of fusing the
    gate into the GEMM epilogue, and the time-embedding MLP keeps its pre-activations;
and records one backward closure per layer.  `backward(dy)` runs the closures in reverse:
  * data gradients of conv / linear layers are `ealdm_conv` with transposed (3x3: flipped) weights;
  * weight gradients are `ealdm_conv_wgrad`, accumulated straight into `param.grad` (fp32, PyTorch layout);
  * the gradient of the fp32 residual stream is fp32 with a bf16 shadow for the GEMMs (same dual-write
    trick as the forward); branch joins (residual connections, the UNet skip concatenation) are folded
    into the `add` / `residual` inputs of the adjoint kernels, so no stand-alone elementwise add runs.
"""
from __future__ import annotations

from typing import Callable, List, Optional

import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .ops import Act, ConvIn
from .unet import Dual, UNetEngine, UNetModel


def _grad2d(p: torch.Tensor) -> torch.Tensor:
    """param.grad as a 2-D [out, rest] fp32 view (allocated as zeros on first use)."""
    if p.grad is None:
        p.grad = torch.zeros_like(p, dtype=torch.float32, memory_format=torch.contiguous_format)
    return p.grad.view(p.shape[0], -1)


def _grad1d(p: torch.Tensor) -> torch.Tensor:
    if p.grad is None:
        p.grad = torch.zeros_like(p, dtype=torch.float32)
    return p.grad


def _pack_dgrad(w: torch.Tensor, dtype) -> torch.Tensor:
    """[O, I, kh, kw] -> [I, (kh', kw', O)] with the taps flipped: the weight matrix of the data gradient.
    Linear / 1x1: [O, I] -> [I, O]."""
    w = w.detach().to(dtype)     # cast first: the permuting copies then move half the bytes
    if w.dim() == 2:
        return w.t().contiguous()
    if w.dim() == 3:
        w = w[..., None]
    i = w.shape[1]
    return w.flip(2, 3).permute(1, 2, 3, 0).reshape(i, -1).contiguous()


class UNetTrainEngine(UNetEngine):
    """Forward + backward on the C ABI.  The packed weights are derived from the parameters at construction:
    build a new engine (UNetModel does so automatically) after every optimizer step."""

    _fused_geglu = False
    _fused_ff = False             # the backward needs the GEGLU pre-activation
    _fold_ln = False              # ... and the LayerNorm outputs / statistics
    _phased_upsample = False      # the backward differentiates the explicit nearest-2x + 3x3 form
    _collapse_xattn = False       # ... and the q / k / v form of the cross-attention
    _ln_epilogue = False          # ... and the LayerNorm inputs as separate tensors

    def __init__(self, m: UNetModel, dtype: torch.dtype):
        super().__init__(m, dtype)
        self.ws = ops.Workspace(self.dev)
        self.tape: List[Callable] = []
        self._attach_params()

    # ---- parameter / dgrad-weight bookkeeping --------------------------------------------------------------
    def _attach_params(self):
        """Attach to every packed layer dict the parameters it came from and the transposed weights."""
        m, dt = self.m, self.dt
        def pd(p):
            base = p._base if (p._base is not None and p.dim() == 2) else None   # [:, :, 0, 0] view of a 1x1 conv
            if base is not None:
                return _pack_dgrad(self._c(base)[:, :, 0, 0], dt)
            return _pack_dgrad(self._c(p), dt)

        # Data-gradient weights.  bf16 (tcgen05): the forward-packed matrix itself, read MN-major by the adjoint
        # mode of ealdm_conv -- nothing is transposed, flipped or copied.  fp32 (SIMT parity mode): the transposed,
        # tap-flipped copy.  `adj` records (tensor, is_adjoint).
        tc = dt == torch.bfloat16

        def adj(fwd, param):
            return (fwd, True) if tc else (pd(param), False)

        def walk(seq, packed):
            for layer, d in zip(seq, packed):
                k = d["kind"]
                d["mod"] = layer
                if k == "res":
                    co = d["cout"]
                    d["wd1"] = adj(d["conv1"].w, layer.in_layers[2].weight)
                    d["wd2"] = adj(d["conv2"].w[:, :9 * co], layer.out_layers[3].weight)
                    if d["skip"]:
                        d["wds"] = adj(d["conv2"].w[:, 9 * co:], layer.skip_connection.weight[:, :, 0, 0])
                elif k == "st":
                    d["wd_pi"] = adj(d["proj_in"].w, layer.proj_in.weight[:, :, 0, 0])
                    d["wd_po"] = adj(d["proj_out"].w, layer.proj_out.weight[:, :, 0, 0])
                    for tb, t in zip(layer.transformer_blocks, d["blocks"]):
                        t["mod"] = tb
                        t["wd_qkv"] = (t["qkv"], True) if tc else (t["qkv"].t().contiguous(), False)
                        t["wd_o1"] = adj(t["o1"].w, tb.attn1.to_out[0].weight)
                        t["wd_q2"] = adj(t["q2"], tb.attn2.to_q.weight)
                        t["wd_o2"] = adj(t["o2"].w, tb.attn2.to_out[0].weight)
                        t["ff1n"] = self._c(tb.ff.net[0].proj.weight)
                        t["ff1n_b"] = tb.ff.net[0].proj.bias.detach().float().contiguous()
                        t["wd_ff1"] = (t["ff1n"], True) if tc else (t["ff1n"].t().contiguous(), False)
                        t["wd_ff2"] = adj(t["ff2"].w, tb.ff.net[2].weight)
                elif k in ("down", "up", "conv_in"):
                    conv = layer.op if k == "down" else (layer.conv if k == "up" else layer)
                    d["wd"] = adj(d["conv"].w, conv.weight) if k != "conv_in" else (pd(conv.weight), False)
                elif k == "ab":
                    raise NotImplementedError("training through AttentionBlock (uncond_cin config) is not built; "
                                              "EALDM trains the SpatialTransformer UNet (stdiff config)")

        for blk, packed in zip(m.input_blocks, self.inp):
            walk(blk, packed)
        walk(m.middle_block, self.mid)
        for blk, packed in zip(m.output_blocks, self.outb):
            walk(blk, packed)
        self.wd_out = (pd(m.out[2].weight), False)          # 4-channel gradient: SIMT kernel, transposed copy
        self.wd_te2 = adj(self.te2.w, m.time_embed[2].weight)
        self.wd_emb = (self.emb_w, True) if tc else (self.emb_w.t().contiguous(), False)
        self.wd_kv = None if self.kv_w is None else ((self.kv_w, True) if tc else (self.kv_w.t().contiguous(), False))

    def _dconv(self, src: Act, wd, out: Act, ksize: int = 3, **kw) -> Act:
        """Data gradient of a 3x3 (pad 1) / 1x1 convolution: conv of the output gradient with the adjoint weights."""
        w, is_adj = wd
        return ops.conv([ConvIn(src, ksize, 1, ksize // 2)], w, out, adjoint=is_adj, **kw)

    def _dlinear(self, src: Act, wd, out: Act, **kw) -> Act:
        w, is_adj = wd
        return ops.linear(src, w, out, adjoint=is_adj, **kw)

    def _stats(self, n):
        return torch.empty((n, 32, 2), dtype=torch.float32, device=self.dev)

    def _gdual(self, n, h, w, c) -> Dual:
        return self._new_dual(n, h, w, c, gn=False)   # gradients need no GroupNorm statistics

    # ---- layers: forward that records its backward -----------------------------------------------------------
    def _res(self, d, x: Dual, emb_all, dest: Optional[Dual]) -> Dual:
        rb = d["mod"]
        n, h, w = x.f.n, x.f.h, x.f.w
        cin, cout = x.f.c, d["cout"]
        st1, st2 = self._stats(n), self._stats(n)
        hn = self._new(n, h, w, cin)
        ops.group_norm(x.f, d["gn1"][0], d["gn1"][1], 1e-5, hn, self.stats, silu=True, stats_out=st1)
        h1 = self._new(n, h, w, cout, torch.float32)
        if self.dt == torch.bfloat16:
            h1.with_gn_partial()
        ops.conv([ConvIn(hn, 3, 1, 1)], d["conv1"].w, h1, bias=d["conv1"].b, rowvec=emb_all, rowvec_col0=d["emb_col0"])
        hn2 = self._new(n, h, w, cout)
        ops.group_norm(h1, d["gn2"][0], d["gn2"][1], 1e-5, hn2, self.stats, silu=True, stats_out=st2)
        out = dest if dest is not None else self._new_dual(n, h, w, cout)
        if d["skip"]:
            ops.conv([ConvIn(hn2, 3, 1, 1), ConvIn(x.h, 1, 1, 0)], d["conv2"].w, out.f, bias=d["conv2"].b,
                     out2=self._out2(out))
        else:
            ops.conv([ConvIn(hn2, 3, 1, 1)], d["conv2"].w, out.f, bias=d["conv2"].b, residual=x.f,
                     out2=self._out2(out))

        def bwd(g: Dual, extra: Optional[Act]) -> Dual:
            ws = self.ws
            ops.colsum(g.h, _grad1d(rb.out_layers[3].bias), ws)
            ops.conv_wgrad(hn2, g.h, _grad2d(rb.out_layers[3].weight), ws, ksize=3, pad=1)
            d_hn2 = self._new(n, h, w, cout)
            self._dconv(g.h, d["wd2"], d_hn2)
            d_h1 = self._new(n, h, w, cout)
            ops.group_norm_bwd(h1, d_hn2, st2, d["gn2"][0], d["gn2"][1], d_h1, ws, silu=True,
                               dgamma=_grad1d(rb.out_layers[0].weight), dbeta=_grad1d(rb.out_layers[0].bias))
            # per-image column sums: gradient of the timestep-embedding row vector (and of conv1's bias)
            ops.colsum(d_h1, self.d_emb_all, ws, segs=n, col0=d["emb_col0"], accumulate=False)
            ops.conv_wgrad(hn, d_h1, _grad2d(rb.in_layers[2].weight), ws, ksize=3, pad=1)
            d_hn = self._new(n, h, w, cin)
            self._dconv(d_h1, d["wd1"], d_hn)
            dx = self._gdual(n, h, w, cin)
            gn1 = dict(silu=True, dgamma=_grad1d(rb.in_layers[0].weight), dbeta=_grad1d(rb.in_layers[0].bias))
            if d["skip"]:
                ops.colsum(g.h, _grad1d(rb.skip_connection.bias), ws)
                ops.conv_wgrad(x.h, g.h, _grad2d(rb.skip_connection.weight), ws, ksize=1)
                tmp = self._new(n, h, w, cin, torch.float32)
                ops.group_norm_bwd(x.f, d_hn, st1, d["gn1"][0], d["gn1"][1], tmp, ws, add=extra, **gn1)
                self._dconv(g.h, d["wds"], dx.f, ksize=1, residual=tmp, out2=self._out2(dx))
            else:
                ops.group_norm_bwd(x.f, d_hn, st1, d["gn1"][0], d["gn1"][1], dx.f, ws, add=g.f, add2=extra,
                                   dx2=self._out2(dx), **gn1)
            return dx

        self.tape.append(bwd)
        return out

    def _st(self, d, x: Dual, kv_all: Optional[Act], n_ctx: int, dest: Optional[Dual]) -> Dual:
        st = d["mod"]
        C_, heads, dh = d["c"], d["heads"], d["dh"]
        n, h, w = x.f.n, x.f.h, x.f.w
        tok = h * w
        f32 = torch.float32
        if kv_all is None:
            raise RuntimeError("SpatialTransformer needs a context tensor")
        stn = self._stats(n)
        xn = self._new(n, h, w, C_)
        ops.group_norm(x.f, d["norm"][0], d["norm"][1], 1e-6, xn, self.stats, silu=False, stats_out=stn)
        t0 = self._new(n, h, w, C_, f32)
        ops.linear(xn, d["proj_in"].w, t0, bias=d["proj_in"].b)
        saved = []
        t = t0
        th = None
        nblk = len(d["blocks"])
        for bi, tb in enumerate(d["blocks"]):
            s = {"t0": t}
            s["a1"] = ops.layer_norm(t, tb["ln1"][0], tb["ln1"][1], 1e-5, self._new(n, h, w, C_))
            s["qkv"] = ops.linear(s["a1"], tb["qkv"], self._new(n, h, w, 3 * C_))
            q = s["qkv"]
            # the tensor-core backward re-creates the probabilities from the forward's log-sum-exp
            s["lse"] = (torch.empty((n, heads, tok), dtype=f32, device=self.dev)
                        if (self.dt == torch.bfloat16 and dh == 32 and tok % 64 == 0) else None)
            s["o"] = ops.attention(q.cols(0, C_), q.cols(C_, C_), q.cols(2 * C_, C_), self._new(n, h, w, C_), batch=n,
                                   heads=heads, head_dim=dh, n_q=tok, n_kv=tok, scale=dh ** -0.5, lse=s["lse"])
            s["t1"] = ops.linear(s["o"], tb["o1"].w, self._new(n, h, w, C_, f32), bias=tb["o1"].b, residual=t)
            s["a2"] = ops.layer_norm(s["t1"], tb["ln2"][0], tb["ln2"][1], 1e-5, self._new(n, h, w, C_))
            s["q2"] = ops.linear(s["a2"], tb["q2"], self._new(n, h, w, C_))
            kc = tb["kv_col0"]
            s["o2"] = ops.attention(s["q2"], kv_all.cols(kc, C_), kv_all.cols(kc + C_, C_), self._new(n, h, w, C_),
                                    batch=n, heads=heads, head_dim=dh, n_q=tok, n_kv=n_ctx, scale=dh ** -0.5)
            s["t2"] = ops.linear(s["o2"], tb["o2"].w, self._new(n, h, w, C_, f32), bias=tb["o2"].b, residual=s["t1"])
            s["a3"] = ops.layer_norm(s["t2"], tb["ln3"][0], tb["ln3"][1], 1e-5, self._new(n, h, w, C_))
            s["pre"] = ops.linear(s["a3"], tb["ff1n"], self._new(n, h, w, 8 * C_), bias=tb["ff1n_b"])
            s["gg"] = ops.geglu(s["pre"], self._new(n, h, w, 4 * C_))
            last = bi == nblk - 1
            t3 = self._new(n, h, w, C_, f32)
            th = self._new(n, h, w, C_) if (last and self.dt != f32) else None
            ops.linear(s["gg"], tb["ff2"].w, t3, bias=tb["ff2"].b, residual=s["t2"], out2=th)
            t = t3
            saved.append(s)
        t_op = th if th is not None else t   # operand of proj_out
        out = dest if dest is not None else self._new_dual(n, h, w, C_)
        ops.linear(t_op, d["proj_out"].w, out.f, bias=d["proj_out"].b, residual=x.f, out2=self._out2(out))

        def bwd(g: Dual, extra: Optional[Act]) -> Dual:
            ws = self.ws
            nd = lambda c, dtype=None: self._new(n, h, w, c, dtype)  # noqa: E731
            ops.colsum(g.h, _grad1d(st.proj_out.bias), ws)
            ops.linear_wgrad(t_op, g.h, _grad2d(st.proj_out.weight), ws)
            dt_ = self._gdual(n, h, w, C_)
            self._dlinear(g.h, d["wd_po"], dt_.f, out2=self._out2(dt_))
            for tb, s in zip(reversed(d["blocks"]), reversed(saved)):
                mod = tb["mod"]
                # feed-forward
                ops.colsum(dt_.h, _grad1d(mod.ff.net[2].bias), ws)
                ops.linear_wgrad(s["gg"], dt_.h, _grad2d(mod.ff.net[2].weight), ws)
                dgg = self._dlinear(dt_.h, tb["wd_ff2"], nd(4 * C_))
                dpre = ops.geglu_bwd(s["pre"], dgg, nd(8 * C_))
                ops.colsum(dpre, _grad1d(mod.ff.net[0].proj.bias), ws)
                ops.linear_wgrad(s["a3"], dpre, _grad2d(mod.ff.net[0].proj.weight), ws)
                da3 = self._dlinear(dpre, tb["wd_ff1"], nd(C_))
                dt2 = self._gdual(n, h, w, C_)
                ops.layer_norm_bwd(s["t2"], da3, tb["ln3"][0], 1e-5, dt2.f, ws, add=dt_.f, dx2=self._out2(dt2),
                                   dgamma=_grad1d(mod.norm3.weight), dbeta=_grad1d(mod.norm3.bias))
                # cross-attention
                ops.colsum(dt2.h, _grad1d(mod.attn2.to_out[0].bias), ws)
                ops.linear_wgrad(s["o2"], dt2.h, _grad2d(mod.attn2.to_out[0].weight), ws)
                do2 = self._dlinear(dt2.h, tb["wd_o2"], nd(C_))
                kc = tb["kv_col0"]
                dq2 = nd(C_)
                ops.attention_bwd(s["q2"], kv_all.cols(kc, C_), kv_all.cols(kc + C_, C_), s["o2"], do2, dq2,
                                  self.d_kv_all.cols(kc, C_), self.d_kv_all.cols(kc + C_, C_), ws, batch=n,
                                  heads=heads, head_dim=dh, n_q=tok, n_kv=n_ctx, scale=dh ** -0.5)
                ops.linear_wgrad(s["a2"], dq2, _grad2d(mod.attn2.to_q.weight), ws)
                da2 = self._dlinear(dq2, tb["wd_q2"], nd(C_))
                dt1 = self._gdual(n, h, w, C_)
                ops.layer_norm_bwd(s["t1"], da2, tb["ln2"][0], 1e-5, dt1.f, ws, add=dt2.f, dx2=self._out2(dt1),
                                   dgamma=_grad1d(mod.norm2.weight), dbeta=_grad1d(mod.norm2.bias))
                # self-attention
                ops.colsum(dt1.h, _grad1d(mod.attn1.to_out[0].bias), ws)
                ops.linear_wgrad(s["o"], dt1.h, _grad2d(mod.attn1.to_out[0].weight), ws)
                do = self._dlinear(dt1.h, tb["wd_o1"], nd(C_))
                q = s["qkv"]
                dqkv = nd(3 * C_)
                ops.attention_bwd(q.cols(0, C_), q.cols(C_, C_), q.cols(2 * C_, C_), s["o"], do, dqkv.cols(0, C_),
                                  dqkv.cols(C_, C_), dqkv.cols(2 * C_, C_), ws, batch=n, heads=heads, head_dim=dh,
                                  n_q=tok, n_kv=tok, scale=dh ** -0.5, lse=s["lse"])
                for j, lin in enumerate((mod.attn1.to_q, mod.attn1.to_k, mod.attn1.to_v)):
                    ops.linear_wgrad(s["a1"], dqkv.cols(j * C_, C_), _grad2d(lin.weight), ws)
                da1 = self._dlinear(dqkv, tb["wd_qkv"], nd(C_))
                dt0 = self._gdual(n, h, w, C_)
                ops.layer_norm_bwd(s["t0"], da1, tb["ln1"][0], 1e-5, dt0.f, ws, add=dt1.f, dx2=self._out2(dt0),
                                   dgamma=_grad1d(mod.norm1.weight), dbeta=_grad1d(mod.norm1.bias))
                dt_ = dt0
            ops.colsum(dt_.h, _grad1d(st.proj_in.bias), ws)
            ops.linear_wgrad(xn, dt_.h, _grad2d(st.proj_in.weight), ws)
            dxn = self._dlinear(dt_.h, d["wd_pi"], nd(C_))
            dx = self._gdual(n, h, w, C_)
            ops.group_norm_bwd(x.f, dxn, stn, d["norm"][0], d["norm"][1], dx.f, ws, silu=False, add=g.f, add2=extra,
                               dx2=self._out2(dx), dgamma=_grad1d(st.norm.weight), dbeta=_grad1d(st.norm.bias))
            return dx

        self.tape.append(bwd)
        return out

    def _ab(self, d, x, dest):
        raise NotImplementedError("training through AttentionBlock is not built")

    def _run(self, layers, x: Dual, emb_all, kv_all, n_ctx, dest: Optional[Dual]) -> Dual:
        for i, d in enumerate(layers):
            dst = dest if i == len(layers) - 1 else None
            k = d["kind"]
            n, h, w = x.f.n, x.f.h, x.f.w
            if k == "res":
                x = self._res(d, x, emb_all, dst)
            elif k == "st":
                x = self._st(d, x, kv_all, n_ctx, dst)
            elif k == "conv_in":
                x = self._conv_in(d, x, dst)
            elif k == "down":
                x = self._down(d, x, dst)
            elif k == "up":
                x = self._up(d, x, dst)
            else:
                raise ValueError(k)
        return x

    def _conv_in(self, d, x: Dual, dst) -> Dual:
        conv = d["mod"]
        n, h, w = x.f.n, x.f.h, x.f.w
        out = dst if dst is not None else self._new_dual(n, h, w, d["conv"].cout)
        self._conv_first(d, x, out)

        def bwd(g: Dual, extra):
            ops.colsum(g.h, _grad1d(conv.bias), self.ws)
            ops.conv_wgrad(x.h, g.h, _grad2d(conv.weight), self.ws, ksize=3, pad=1)
            if not self.need_dx:
                return None
            dx = self._new(n, h, w, x.h.c, torch.float32)
            self._dconv(g.h, d["wd"], dx)
            return Dual(dx, dx)

        self.tape.append(bwd)
        return out

    def _down(self, d, x: Dual, dst) -> Dual:
        conv = d["mod"].op
        n, h, w = x.f.n, x.f.h, x.f.w
        out = dst if dst is not None else self._new_dual(n, h // 2, w // 2, d["conv"].cout)
        ops.conv([ConvIn(x.h, 3, 2, 1)], d["conv"].w, out.f, bias=d["conv"].b, out2=self._out2(out))

        def bwd(g: Dual, extra):
            ops.colsum(g.h, _grad1d(conv.bias), self.ws)
            ops.conv_wgrad(x.h, g.h, _grad2d(conv.weight), self.ws, ksize=3, stride=2, pad=1)
            z = ops.zero_insert2x(g.h, self._new(n, h, w, g.h.c))
            dx = self._gdual(n, h, w, x.f.c)
            self._dconv(z, d["wd"], dx.f, residual=extra, out2=self._out2(dx))
            return dx

        self.tape.append(bwd)
        return out

    def _up(self, d, x: Dual, dst) -> Dual:
        conv = d["mod"].conv
        n, h, w = x.f.n, x.f.h, x.f.w
        out = dst if dst is not None else self._new_dual(n, h * 2, w * 2, d["conv"].cout)
        up = ops.upsample_nearest2x(x.h, self._new(n, h * 2, w * 2, x.h.c))
        ops.conv([ConvIn(up, 3, 1, 1)], d["conv"].w, out.f, bias=d["conv"].b, out2=self._out2(out))

        def bwd(g: Dual, extra):
            ops.colsum(g.h, _grad1d(conv.bias), self.ws)
            ops.conv_wgrad(up, g.h, _grad2d(conv.weight), self.ws, ksize=3, pad=1)
            dup = self._new(n, 2 * h, 2 * w, x.h.c)
            self._dconv(g.h, d["wd"], dup)
            dx = self._gdual(n, h, w, x.f.c)
            ops.sumpool2x2(dup, dx.f, add=extra, dx2=self._out2(dx))
            return dx

        self.tape.append(bwd)
        return out

    # ---- forward --------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x: torch.Tensor, timesteps: torch.Tensor, context: Optional[torch.Tensor]) -> torch.Tensor:
        m, dt, dev = self.m, self.dt, self.dev
        f32 = torch.float32
        n, cin, H, W = x.shape
        assert cin == m.in_channels
        x = x.float().contiguous()
        timesteps = timesteps.to(device=dev, dtype=torch.int64).contiguous()
        self.tape = []
        self.stats = ops.group_norm_workspace(n, H * W, 2 * self.mid_ch, dev)

        # time embedding MLP with saved pre-activations
        temb = Act.empty(1, 1, n, m.model_channels, dt, dev)
        ops.timestep_embedding(timesteps, m.model_channels, self.freqs, temb.buf)
        ted = self.te0.w.shape[0]
        e0 = ops.linear(temb, self.te0.w, Act.empty(1, 1, n, ted, f32, dev), bias=self.te0.b)
        e1 = ops.silu(e0, Act.empty(1, 1, n, ted, dt, dev))
        e2 = ops.linear(e1, self.te2.w, Act.empty(1, 1, n, ted, f32, dev), bias=self.te2.b)
        semb = ops.silu(e2, Act.empty(1, 1, n, ted, dt, dev))
        emb_all = ops.linear(semb, self.emb_w, Act.empty(1, 1, n, self.emb_cols, f32, dev), bias=self.emb_b)
        self.d_emb_all = torch.empty((n, self.emb_cols), dtype=f32, device=dev)

        kv_all, n_ctx, ctx = None, 0, None
        if self.kv_w is not None:
            if context is None:
                raise RuntimeError("this UNet was built with context_dim: pass context=[N, T, context_dim]")
            n_ctx = context.shape[1]
            csrc = Act(context.detach().float().reshape(n * n_ctx, -1).contiguous(), n, 1, n_ctx)
            ctx = ops.copy2d(csrc, Act.empty(n, 1, n_ctx, csrc.c, dt, dev)) if dt != f32 else csrc
            kv_all = ops.linear(ctx, self.kv_w, Act.empty(n, 1, n_ctx, self.kv_cols, dt, dev))
            self.d_kv_all = Act.empty(n, 1, n_ctx, self.kv_cols, dt, dev)

        n_in = len(self.inp)
        res = [(H, W)]
        for layers in self.inp[1:]:
            hh, ww = res[-1]
            res.append((hh // 2, ww // 2) if layers[0]["kind"] == "down" else (hh, ww))
        cat: List[Dual] = []
        h_ch = self.mid_ch
        h_chs = []
        for j, layers in enumerate(self.outb):
            i = n_in - 1 - j
            hh, ww = res[i]
            cat.append(self._new_dual(n, hh, ww, h_ch + self.skip_ch[i]))
            h_chs.append(h_ch)
            h_ch = layers[0]["cout"]

        def window(dl: Dual, c0, c):
            return Dual(dl.f.cols(c0, c), dl.f.cols(c0, c) if dl.h is dl.f else dl.h.cols(c0, c))

        skip_dst = [window(cat[n_in - 1 - i], cat[n_in - 1 - i].f.c - self.skip_ch[i], self.skip_ch[i])
                    for i in range(n_in)]
        # the first conv reads a 4-channel tensor: pad its pitch to 8 so that the rows are 16-byte aligned
        xin = Act(torch.zeros((n * H * W, 8), dtype=dt, device=dev), n, H, W, cin, 0)
        ops.nchw_to_nhwc(x, xin)
        h = Dual(xin, xin)
        marks = []   # tape index ranges per UNet block, in forward order
        for i, layers in enumerate(self.inp):
            t0 = len(self.tape)
            h = self._run(layers, h, emb_all.buf, kv_all, n_ctx, skip_dst[i])
            marks.append(("in", i, t0, len(self.tape)))
        t0 = len(self.tape)
        h = self._run(self.mid, h, emb_all.buf, kv_all, n_ctx, window(cat[0], 0, self.mid_ch))
        marks.append(("mid", 0, t0, len(self.tape)))
        for j, layers in enumerate(self.outb):
            dst = window(cat[j + 1], 0, h_chs[j + 1]) if j + 1 < len(self.outb) else None
            t0 = len(self.tape)
            h = self._run(layers, cat[j], emb_all.buf, kv_all, n_ctx, dst)
            marks.append(("out", j, t0, len(self.tape)))

        st_out = self._stats(n)
        hn = self._new(n, H, W, h.f.c)
        ops.group_norm(h.f, self.out_norm[0], self.out_norm[1], 1e-5, hn, self.stats, silu=True, stats_out=st_out)
        co = m.out_channels
        co_pad = (co + 7) // 8 * 8
        obuf = Act.empty(n, H, W, co_pad, f32, dev)
        ops.conv([ConvIn(hn, 3, 1, 1)], self.out_conv.w, obuf.cols(0, co), bias=self.out_conv.b)
        y = torch.empty((n, co, H, W), dtype=f32, device=dev)
        ops.nhwc_to_nchw(obuf.cols(0, co), y)

        self._saved = dict(n=n, H=H, W=W, co=co, co_pad=co_pad, hn=hn, h_last=h, st_out=st_out, marks=marks,
                           h_chs=h_chs, n_in=n_in, temb=temb, e0=e0, e1=e1, e2=e2, semb=semb, ctx=ctx, n_ctx=n_ctx,
                           cin=cin)
        return y

    # ---- backward -------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def backward(self, dy: torch.Tensor, need_dx: bool = False, need_dcontext: bool = False):
        """Accumulate d(loss)/d(param) into every parameter's .grad given dy = d(loss)/d(output) [N, C, H, W];
        returns (dx or None, dcontext or None)."""
        s, m, dt, dev, ws = self._saved, self.m, self.dt, self.dev, self.ws
        f32 = torch.float32
        n, H, W, co = s["n"], s["H"], s["W"], s["co"]
        self.need_dx = need_dx
        dy = dy.detach().float().contiguous()
        assert tuple(dy.shape) == (n, co, H, W)
        # output head: GN -> SiLU -> conv3x3
        dyo = Act(torch.zeros((n * H * W, s["co_pad"]), dtype=dt, device=dev), n, H, W, co, 0)
        ops.nchw_to_nhwc(dy, dyo)
        ops.colsum(dyo, _grad1d(m.out[2].bias), ws)
        ops.conv_wgrad(s["hn"], dyo, _grad2d(m.out[2].weight), ws, ksize=3, pad=1)
        dhn = self._new(n, H, W, s["hn"].c)
        self._dconv(dyo, self.wd_out, dhn)
        h_last: Dual = s["h_last"]
        g = self._gdual(n, H, W, h_last.f.c)
        ops.group_norm_bwd(h_last.f, dhn, s["st_out"], self.out_norm[0], self.out_norm[1], g.f, ws, silu=True,
                           dx2=self._out2(g), dgamma=_grad1d(m.out[0].weight), dbeta=_grad1d(m.out[0].bias))
        hook = getattr(m, "grad_ready_hook", None)   # parallel.GradBuckets: overlap the all-reduce with the backward
        if hook is not None:
            hook("head", 0)

        def window(dl: Dual, c0, c):
            return Dual(dl.f.cols(c0, c), dl.f.cols(c0, c) if dl.h is dl.f else dl.h.cols(c0, c))

        def run_block(t0, t1, g, extra):
            for ti in range(t1 - 1, t0 - 1, -1):
                g = self.tape[ti](g, extra if ti == t0 else None)
            return g

        n_in = s["n_in"]
        dskip = [None] * n_in
        for kind, idx, t0, t1 in reversed(s["marks"]):
            if kind == "out":
                dcat = run_block(t0, t1, g, None)
                hc = s["h_chs"][idx]
                i = n_in - 1 - idx
                dskip[i] = dcat.f.cols(hc, dcat.f.c - hc)
                g = window(dcat, 0, hc)
            elif kind == "mid":
                g = run_block(t0, t1, g, dskip[n_in - 1])
            else:
                g = run_block(t0, t1, g, dskip[idx - 1] if idx > 0 else None)
            if hook is not None:
                hook(kind, idx)
        dx = None
        if need_dx and g is not None:
            dx = torch.empty((n, s["cin"], H, W), dtype=f32, device=dev)
            ops.nhwc_to_nchw(g.f, dx)

        # timestep-embedding MLP: emb_all = semb W_emb^T + b (all ResBlocks stacked), semb = silu(e2), ...
        d_emb = Act(self.d_emb_all, 1, 1, n)
        d_emb_h = d_emb if dt == f32 else ops.copy2d(d_emb, Act.empty(1, 1, n, self.emb_cols, dt, dev))

        def res_blocks():
            for blk, packed in list(zip(m.input_blocks, self.inp)) + [(m.middle_block, self.mid)] + \
                    list(zip(m.output_blocks, self.outb)):
                for layer, d in zip(blk, packed):
                    if d["kind"] == "res":
                        yield layer, d

        for rb, d in res_blocks():
            c0, cc = d["emb_col0"], d["cout"]
            ops.linear_wgrad(s["semb"], d_emb_h.cols(c0, cc), _grad2d(rb.emb_layers[1].weight), ws)
            ops.colsum(d_emb.cols(c0, cc), _grad1d(rb.emb_layers[1].bias), ws)
            ops.colsum(d_emb.cols(c0, cc), _grad1d(rb.in_layers[2].bias), ws)
        ted = self.te0.w.shape[0]
        d_semb = self._dlinear(d_emb_h, self.wd_emb, Act.empty(1, 1, n, ted, dt, dev))
        d_e2 = ops.silu_bwd(s["e2"], d_semb, Act.empty(1, 1, n, ted, dt, dev))
        ops.linear_wgrad(s["e1"], d_e2, _grad2d(m.time_embed[2].weight), ws)
        ops.colsum(d_e2, _grad1d(m.time_embed[2].bias), ws)
        d_e1 = self._dlinear(d_e2, self.wd_te2, Act.empty(1, 1, n, ted, dt, dev))
        d_e0 = ops.silu_bwd(s["e0"], d_e1, Act.empty(1, 1, n, ted, dt, dev))
        ops.linear_wgrad(s["temb"], d_e0, _grad2d(m.time_embed[0].weight), ws)
        ops.colsum(d_e0, _grad1d(m.time_embed[0].bias), ws)

        # context K/V projections of every cross-attention layer
        dcontext = None
        if self.kv_w is not None:
            for blk, packed in list(zip(m.input_blocks, self.inp)) + [(m.middle_block, self.mid)] + \
                    list(zip(m.output_blocks, self.outb)):
                for layer, d in zip(blk, packed):
                    if d["kind"] != "st":
                        continue
                    for tb in d["blocks"]:
                        kc, C_ = tb["kv_col0"], d["c"]
                        ops.linear_wgrad(s["ctx"], self.d_kv_all.cols(kc, C_), _grad2d(tb["mod"].attn2.to_k.weight), ws)
                        ops.linear_wgrad(s["ctx"], self.d_kv_all.cols(kc + C_, C_),
                                         _grad2d(tb["mod"].attn2.to_v.weight), ws)
            if need_dcontext:
                dc = self._dlinear(self.d_kv_all, self.wd_kv, Act.empty(n, 1, s["n_ctx"], self.kv_w.shape[1], f32, dev))
                dcontext = dc.buf.reshape(n, s["n_ctx"], -1)
        self.tape = []
        self._saved = None
        return dx, dcontext


class _UNetFunction(torch.autograd.Function):
    """Bridges the engine into torch.autograd: parameter gradients are accumulated into .grad by the engine
    itself (they are not autograd inputs); `anchor` is any parameter, passed only so that autograd calls us."""

    @staticmethod
    def forward(ctx, anchor, x, context, timesteps, engine: UNetTrainEngine):
        ctx.engine = engine
        ctx.need_dx = x.requires_grad
        ctx.need_dc = context is not None and context.requires_grad
        return engine.forward(x, timesteps, context)

    @staticmethod
    def backward(ctx, dy):
        dx, dc = ctx.engine.backward(dy, need_dx=ctx.need_dx, need_dcontext=ctx.need_dc)
        return None, dx, dc, None, None


def unet_forward_train(model: UNetModel, x, timesteps, context):
    """Differentiable forward used by UNetModel.forward when autograd is recording."""
    engine = UNetTrainEngine(model, model._compute_dtype)
    anchor = next(p for p in model.parameters() if p.requires_grad)
    return _UNetFunction.apply(anchor, x, context, timesteps, engine)


class FusedTrainStep:
    """One optimisation step's forward + backward of `LatentDiffusion.p_losses` (ddpm.py:1036-1078) as a straight
    sequence of libealdm_b200 launches -- q_sample, UNet forward at 2B, guided-eps MSE, its adjoint, UNet backward,
    including the per-step weight packing -- with no autograd graph in between, so that the WHOLE step can be
    captured once and replayed as one CUDA graph (~1800 launches; eager Python issue time would otherwise be a
    third of the step).  Gradients accumulate into the flat buffer of `parallel.GradBuckets` (created here);
    call `buckets.zero_()` before and `buckets.finish()` after the step (finish() runs the data-parallel
    all-reduce: overlapped with the backward in eager mode, after the replay in graph mode)."""

    def __init__(self, ld, use_graph: bool = True, bucket_mb: float = 64.0, group=None):
        from .parallel import GradBuckets
        self.ld = ld
        self.unet: UNetModel = ld.model.diffusion_model
        self.use_graph = use_graph
        self.buckets = GradBuckets(self.unet, bucket_mb=bucket_mb, group=group)
        self._graphs = {}
        self.force_segments = False      # tests: cut the capture into segments even on one rank
        dev = ld.betas.device
        self._logvar = ld.logvar.detach().to(dev).float()

    def _run(self, x0, c2, t, noise):
        ld, unet = self.ld, self.unet
        B = x0.shape[0]
        engine = UNetTrainEngine(unet, unet._compute_dtype)
        x_noisy = ops.q_sample(x0, noise, t, ld.sqrt_alphas_cumprod, ld.sqrt_one_minus_alphas_cumprod)
        target = noise if ld.parameterization == "eps" else x0
        s = float(ld.unconditional_guidance_scale)
        if s != 1.0:
            eps = engine.forward(torch.cat([x_noisy] * 2), torch.cat([t] * 2), c2)
            e_u, e_c = eps[:B], eps[B:]
        else:
            e_u, e_c = None, engine.forward(x_noisy, t, c2)
        loss_simple = ops.cfg_mse(e_c, target, e_uncond=e_u, cfg_scale=s)
        logvar_t = self._logvar[t]
        lvlb_t = ld.lvlb_weights[t]
        loss = ld.l_simple_weight * (loss_simple / torch.exp(logvar_t) + logvar_t).mean() \
            + ld.original_elbo_weight * (lvlb_t * loss_simple).mean()
        w = ((ld.l_simple_weight / torch.exp(logvar_t) + ld.original_elbo_weight * lvlb_t) / B).float().contiguous()
        de_u, de_c = ops.cfg_mse_bwd(e_c, target, w, e_uncond=e_u, cfg_scale=s)
        engine.backward(de_c if de_u is None else torch.cat([de_u, de_c]))
        return loss

    @torch.no_grad()
    def __call__(self, x0, cond, t, noise):
        """x0 [B,4,H,W], cond [2B,T,D] = cat([c_neg, c]) (or [B,T,D] without guidance), t [B] int64, noise like x0.
        Returns the loss (a device tensor); parameter gradients are accumulated."""
        x0, noise = x0.float().contiguous(), noise.float().contiguous()
        cond, t = cond.float().contiguous(), t.to(torch.int64).contiguous()
        if not self.use_graph:
            return self._run(x0, cond, t, noise)
        key = (tuple(x0.shape), tuple(cond.shape), self.unet._compute_dtype)
        ent = self._graphs.get(key)
        if ent is None:
            ent = self._capture(x0.clone(), cond.clone(), t.clone(), noise.clone())
            self._graphs[key] = ent
        segments, sx, sc, st, sn, loss = ent
        sx.copy_(x0); sc.copy_(cond); st.copy_(t); sn.copy_(noise)
        for graph, nready in segments:
            graph.replay()
            if nready is not None:        # the gradients [0, nready) of the flat buffer are final: reduce them now
                self.buckets._launch_upto(nready)
        return loss

    def _capture(self, sx, sc, st, sn):
        """Capture the step as CUDA graphs.  With one rank it is a single graph.  With data parallelism the capture
        is cut at `segment_marks` (UNet blocks whose backward has just finished): between two replays the host
        launches the NCCL all-reduce of the gradient buckets that are complete, so the exchange overlaps the rest of
        the backward exactly as in eager mode, while the ~1800 kernel launches still cost no host time."""
        unet, gb = self.unet, self.buckets
        hook, unet.grad_ready_hook = unet.grad_ready_hook, None
        keep = gb.flat.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):     # eager warm-up: lazy kernel attributes, allocator, workspace growth
            self._run(sx, sc, st, sn)
        torch.cuda.current_stream().wait_stream(side)
        marks = set(self.segment_marks) if (gb.world > 1 or self.force_segments) else set()
        segments = []
        state = {"graph": None}
        pool = torch.cuda.graph_pool_handle()
        cap_stream = torch.cuda.Stream()

        def begin():
            g = torch.cuda.CUDAGraph()
            g.capture_begin(pool=pool)
            state["graph"] = g

        def end(nready):
            state["graph"].capture_end()
            segments.append((state["graph"], nready))

        def cut(kind, idx):
            if (kind, idx) in marks:
                end(gb.ready_at[(kind, idx)])
                begin()

        unet.grad_ready_hook = cut
        cap_stream.wait_stream(torch.cuda.current_stream())
        torch.cuda.synchronize()
        with torch.cuda.stream(cap_stream):
            begin()
            loss = self._run(sx, sc, st, sn)
            end(None)
        torch.cuda.current_stream().wait_stream(cap_stream)
        gb.flat.copy_(keep)               # the warm-up accumulated one extra gradient: undo it
        unet.grad_ready_hook = hook
        return segments, sx, sc, st, sn, loss

    # UNet blocks after whose backward the capture is cut (kind, index) -- three exchanges overlap the backward
    segment_marks = (("out", 6), ("out", 3), ("mid", 0), ("in", 4))
