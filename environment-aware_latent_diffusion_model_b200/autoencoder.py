"""Drop-in `AutoencoderKL` first stage (reference: ldm/models/autoencoder.py:285-333 with
Encoder / Decoder / ResnetBlock / AttnBlock / Up- / Downsample from
ldm/modules/diffusionmodules/model.py:33-202,368-568 and DiagonalGaussianDistribution from
ldm/modules/distributions/distributions.py:24-62).

As in unet.py the module tree only owns the parameters under the reference's names; `encode` /
`decode` run on `AutoencoderEngine`, which reuses the same libealdm_b200 kernels as the UNet
(GroupNorm eps 1e-6 + swish, implicit-GEMM 3x3 convolutions with the 1x1 nin_shortcut accumulated in
the same GEMM, asymmetric-pad stride-2 downsampling, nearest-2x upsampling).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .ops import Act, ConvIn
from .packing import pack_conv_weight, pack_upsample_phases
from .unet import Dual
from .util import instantiate_from_config


class DiagonalGaussianDistribution(object):
    """distributions.py:24-62"""

    def __init__(self, parameters, deterministic=False):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.deterministic = deterministic
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)
        if self.deterministic:
            self.var = self.std = torch.zeros_like(self.mean)

    def sample(self):
        # drawn on the CPU then moved, like the reference (distributions.py:35-37), to keep RNG parity
        return self.mean + self.std * torch.randn(self.mean.shape).to(device=self.parameters.device)

    def kl(self, other=None):
        if self.deterministic:
            return torch.Tensor([0.])
        if other is None:
            return 0.5 * torch.sum(torch.pow(self.mean, 2) + self.var - 1.0 - self.logvar, dim=[1, 2, 3])
        return 0.5 * torch.sum(torch.pow(self.mean - other.mean, 2) / other.var + self.var / other.var
                               - 1.0 - self.logvar + other.logvar, dim=[1, 2, 3])

    def nll(self, sample, dims=(1, 2, 3)):
        """distributions.py:54-60"""
        if self.deterministic:
            return torch.Tensor([0.])
        import math
        return 0.5 * torch.sum(math.log(2.0 * math.pi) + self.logvar + torch.pow(sample - self.mean, 2) / self.var,
                               dim=list(dims))

    def mode(self):
        return self.mean


_posterior_classes = {}


def posterior_class():
    """The class `AutoencoderKL.encode` instantiates.  The reference decides by `isinstance(encoder_posterior,
    DiagonalGaussianDistribution)` (ldm/models/diffusion/ddpm.py:550-557) and raises otherwise, so when this first
    stage is hosted by the reference's LatentDiffusion -- i.e. the reference's distributions module is loaded in this
    process -- the returned object must also be an instance of THAT class."""
    import sys
    ref = sys.modules.get("ldm.modules.distributions.distributions")
    ref_cls = getattr(ref, "DiagonalGaussianDistribution", None) if ref is not None else None
    if ref_cls is None or ref_cls is DiagonalGaussianDistribution:
        return DiagonalGaussianDistribution
    cls = _posterior_classes.get(ref_cls)
    if cls is None:
        cls = type("DiagonalGaussianDistribution", (DiagonalGaussianDistribution, ref_cls), {})
        _posterior_classes[ref_cls] = cls
    return cls


def Normalize(c):  # model.py:38-39
    return nn.GroupNorm(32, c, eps=1e-6, affine=True)


class ResnetBlock(nn.Module):
    """model.py:82-141 (temb_channels = 0 in the autoencoder)"""

    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout=0.0, temb_channels=0):
        super().__init__()
        out_channels = in_channels if out_channels is None else out_channels
        self.in_channels, self.out_channels = in_channels, out_channels
        if conv_shortcut or temb_channels > 0:
            raise NotImplementedError("conv_shortcut / temb are not used by the autoencoder configs")
        self.norm1 = Normalize(in_channels)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.norm2 = Normalize(out_channels)
        self.dropout = nn.Dropout(dropout)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        if in_channels != out_channels:
            self.nin_shortcut = nn.Conv2d(in_channels, out_channels, 1)


class AttnBlock(nn.Module):
    """model.py:150-202"""

    def __init__(self, in_channels):
        super().__init__()
        self.in_channels = in_channels
        self.norm = Normalize(in_channels)
        self.q = nn.Conv2d(in_channels, in_channels, 1)
        self.k = nn.Conv2d(in_channels, in_channels, 1)
        self.v = nn.Conv2d(in_channels, in_channels, 1)
        self.proj_out = nn.Conv2d(in_channels, in_channels, 1)


class _Resample(nn.Module):
    def __init__(self, in_channels, with_conv, stride):
        super().__init__()
        if not with_conv:
            raise NotImplementedError("resamp_with_conv=False")
        self.with_conv = with_conv
        self.conv = nn.Conv2d(in_channels, in_channels, 3, stride=stride, padding=0 if stride == 2 else 1)


class Upsample(_Resample):
    """model.py:42-57"""

    def __init__(self, in_channels, with_conv):
        super().__init__(in_channels, with_conv, 1)


class Downsample(_Resample):
    """model.py:60-79: F.pad(x, (0,1,0,1)) then conv stride 2 padding 0"""

    def __init__(self, in_channels, with_conv):
        super().__init__(in_channels, with_conv, 2)


def _check_attn_type(attn_type, use_linear_attn):
    if use_linear_attn or attn_type != "vanilla":
        raise NotImplementedError("only attn_type='vanilla' is supported")


class Encoder(nn.Module):
    """model.py:368-459"""

    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, double_z=True,
                 use_linear_attn=False, attn_type="vanilla", **ignore_kwargs):
        super().__init__()
        _check_attn_type(attn_type, use_linear_attn)
        self.ch, self.num_resolutions, self.num_res_blocks = ch, len(ch_mult), num_res_blocks
        self.resolution, self.in_channels = resolution, in_channels
        self.conv_in = nn.Conv2d(in_channels, ch, 3, padding=1)
        curr_res = resolution
        in_ch_mult = (1,) + tuple(ch_mult)
        self.down = nn.ModuleList()
        block_in = ch
        for i_level in range(self.num_resolutions):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_in, block_out = ch * in_ch_mult[i_level], ch * ch_mult[i_level]
            for _ in range(num_res_blocks):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, dropout=dropout))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(AttnBlock(block_in))
            down = nn.Module()
            down.block, down.attn = block, attn
            if i_level != self.num_resolutions - 1:
                down.downsample = Downsample(block_in, resamp_with_conv)
                curr_res //= 2
            self.down.append(down)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, dropout=dropout)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, dropout=dropout)
        self.norm_out = Normalize(block_in)
        self.conv_out = nn.Conv2d(block_in, 2 * z_channels if double_z else z_channels, 3, padding=1)


class Decoder(nn.Module):
    """model.py:462-568"""

    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, give_pre_end=False, tanh_out=False,
                 use_linear_attn=False, attn_type="vanilla", **ignorekwargs):
        super().__init__()
        _check_attn_type(attn_type, use_linear_attn)
        if give_pre_end or tanh_out:
            raise NotImplementedError("give_pre_end / tanh_out")
        self.ch, self.num_resolutions, self.num_res_blocks = ch, len(ch_mult), num_res_blocks
        self.resolution, self.in_channels = resolution, in_channels
        block_in = ch * ch_mult[self.num_resolutions - 1]
        curr_res = resolution // 2 ** (self.num_resolutions - 1)
        self.z_shape = (1, z_channels, curr_res, curr_res)
        self.conv_in = nn.Conv2d(z_channels, block_in, 3, padding=1)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, dropout=dropout)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, dropout=dropout)
        self.up = nn.ModuleList()
        for i_level in reversed(range(self.num_resolutions)):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_out = ch * ch_mult[i_level]
            for _ in range(num_res_blocks + 1):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, dropout=dropout))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(AttnBlock(block_in))
            up = nn.Module()
            up.block, up.attn = block, attn
            if i_level != 0:
                up.upsample = Upsample(block_in, resamp_with_conv)
                curr_res *= 2
            self.up.insert(0, up)
        self.norm_out = Normalize(block_in)
        self.conv_out = nn.Conv2d(block_in, out_ch, 3, padding=1)


class AutoencoderKL(nn.Module):
    """autoencoder.py:285-333.  `compute_dtype` ("bf16" | "fp32") is the only extra keyword."""

    def __init__(self, ddconfig, lossconfig=None, embed_dim=4, ckpt_path=None, ignore_keys=(), image_key="image",
                 colorize_nlabels=None, monitor=None, compute_dtype="bf16"):
        super().__init__()
        self.image_key = image_key
        self.encoder = Encoder(**ddconfig)
        self.decoder = Decoder(**ddconfig)
        self.loss = instantiate_from_config(lossconfig) if lossconfig else nn.Identity()
        assert ddconfig["double_z"]
        self.quant_conv = nn.Conv2d(2 * ddconfig["z_channels"], 2 * embed_dim, 1)
        self.post_quant_conv = nn.Conv2d(embed_dim, ddconfig["z_channels"], 1)
        self.embed_dim = embed_dim
        if monitor is not None:
            self.monitor = monitor
        self._compute_dtype = {"bf16": torch.bfloat16, "fp32": torch.float32}[compute_dtype]
        self._engine: Optional["AutoencoderEngine"] = None
        self.register_load_state_dict_post_hook(lambda m, keys: m.invalidate_packed())
        if ckpt_path is not None:
            sd = torch.load(ckpt_path, map_location="cpu")["state_dict"]
            for k in list(sd.keys()):
                if any(k.startswith(ik) for ik in ignore_keys):
                    del sd[k]
            self.load_state_dict(sd, strict=False)

    def set_compute_dtype(self, name):
        self._compute_dtype = {"bf16": torch.bfloat16, "fp32": torch.float32}[name]
        self._engine = None
        return self

    def invalidate_packed(self):
        self._engine = None

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def _eng(self, x):
        if not x.is_cuda:
            raise RuntimeError("ealdm_b200.AutoencoderKL runs on CUDA (sm_100a) only; there is no CPU fallback")
        # the engine caches kernel-layout weights: compare the parameters' version counters (see unet.UNetModel)
        if self._engine is not None and sum(p._version for p in self._plist) != self._engine_fp:
            self._engine = None
        if self._engine is None:
            self._plist = list(self.parameters())
            self._engine = AutoencoderEngine(self, self._compute_dtype)
            self._engine_fp = sum(p._version for p in self._plist)
        return self._engine

    @torch.no_grad()
    def encode(self, x):
        return posterior_class()(self._eng(x).encode_moments(x))

    @torch.no_grad()
    def decode(self, z):
        return self._eng(z).decode(z)

    def forward(self, input, sample_posterior=True):
        posterior = self.encode(input)
        z = posterior.sample() if sample_posterior else posterior.mode()
        return self.decode(z), posterior


class VectorQuantizer(nn.Module):
    """taming-transformers `VectorQuantizer2` as the reference instantiates it (autoencoder.py:39-41: beta 0.25,
    remap None, sane_index_shape False, legacy True): owns `embedding.weight` [n_e, e_dim]; forward returns
    (z_q, loss, (perplexity, min_encodings, min_encoding_indices)) with the nearest-code search in one kernel.
    The commitment loss is a training quantity of the (frozen) first stage and is returned as None."""

    def __init__(self, n_e, e_dim, beta=0.25, remap=None, unknown_index="random", sane_index_shape=False, legacy=True):
        super().__init__()
        if remap is not None:
            raise NotImplementedError("VectorQuantizer remap")
        self.n_e, self.e_dim, self.beta, self.legacy = n_e, e_dim, beta, legacy
        self.sane_index_shape = sane_index_shape
        self.embedding = nn.Embedding(n_e, e_dim)
        self.embedding.weight.data.uniform_(-1.0 / n_e, 1.0 / n_e)

    @torch.no_grad()
    def forward(self, z, temp=None, rescale_logits=False, return_logits=False):
        if not z.is_cuda:
            raise RuntimeError("ealdm_b200.VectorQuantizer runs on CUDA (sm_100a) only; there is no CPU fallback")
        zq, idx = ops.vq_nearest(z.float().contiguous(), self.embedding.weight.detach().float().contiguous())
        if self.sane_index_shape:
            idx = idx.reshape(z.shape[0], z.shape[2], z.shape[3])
        return zq, None, (None, None, idx)

    def get_codebook_entry(self, indices, shape):
        z_q = self.embedding(indices)
        if shape is not None:
            z_q = z_q.view(shape).permute(0, 3, 1, 2).contiguous()
        return z_q


class VQModelInterface(AutoencoderKL):
    """VQModelInterface / VQModel (autoencoder.py:14-110,263-282), the first stage of the shipped EALDM configs
    (stdiff_cin-ldm-vq-f8.yaml:37-60): `encode` returns the pre-quantisation latent, `decode` quantises (unless
    force_not_quantize), applies post_quant_conv and the decoder.  State-dict names are the reference's
    (`encoder.*`, `decoder.*`, `quantize.embedding.weight`, `quant_conv.*`, `post_quant_conv.*`)."""

    def __init__(self, embed_dim, ddconfig=None, lossconfig=None, n_embed=None, ckpt_path=None, ignore_keys=(),
                 image_key="image", colorize_nlabels=None, monitor=None, batch_resize_range=None,
                 scheduler_config=None, lr_g_factor=1.0, remap=None, sane_index_shape=False, use_ema=False,
                 compute_dtype="bf16"):
        nn.Module.__init__(self)
        if use_ema:
            raise NotImplementedError("EMA of the first stage (training of the autoencoder is out of scope)")
        self.embed_dim, self.n_embed, self.image_key = embed_dim, n_embed, image_key
        self.encoder = Encoder(**ddconfig)
        self.decoder = Decoder(**ddconfig)
        self.loss = instantiate_from_config(lossconfig) if lossconfig else nn.Identity()
        self.quantize = VectorQuantizer(n_embed, embed_dim, beta=0.25, remap=remap, sane_index_shape=sane_index_shape)
        self.quant_conv = nn.Conv2d(ddconfig["z_channels"], embed_dim, 1)
        self.post_quant_conv = nn.Conv2d(embed_dim, ddconfig["z_channels"], 1)
        if monitor is not None:
            self.monitor = monitor
        self._compute_dtype = {"bf16": torch.bfloat16, "fp32": torch.float32}[compute_dtype]
        self._engine = None
        self.register_load_state_dict_post_hook(lambda m, keys: m.invalidate_packed())
        if ckpt_path is not None:
            sd = torch.load(ckpt_path, map_location="cpu")["state_dict"]
            for k in list(sd.keys()):
                if any(k.startswith(ik) for ik in ignore_keys):
                    del sd[k]
            self.load_state_dict(sd, strict=False)

    @torch.no_grad()
    def encode(self, x):
        return self._eng(x).encode_moments(x)       # encoder + quant_conv: the continuous latent h

    @torch.no_grad()
    def decode(self, h, force_not_quantize=False):
        quant = h if force_not_quantize else self.quantize(h)[0]
        return self._eng(h).decode(quant)

    def forward(self, input, return_pred_indices=False):
        h = self.encode(input)
        quant, _, (_, _, ind) = self.quantize(h)
        dec = self._eng(h).decode(quant)
        return (dec, None, ind) if return_pred_indices else (dec, None)


# ---- execution engine ------------------------------------------------------------------------------------
class AutoencoderEngine:
    def __init__(self, m: AutoencoderKL, dtype: torch.dtype):
        L.load()
        self.m, self.dt = m, dtype
        self.dev = next(m.parameters()).device
        assert self.dev.type == "cuda"
        f32 = lambda t: t.detach().float().contiguous()  # noqa: E731
        self._cin_pad = 64 if dtype == torch.bfloat16 else 4

        def pc(conv, pad_cin_to=None):
            w = conv.weight.detach()
            if pad_cin_to is not None and w.shape[1] < pad_cin_to:
                w = torch.cat([w, w.new_zeros(w.shape[0], pad_cin_to - w.shape[1], *w.shape[2:])], dim=1)
            return pack_conv_weight(w, dtype), f32(conv.bias)

        def res(rb: ResnetBlock):
            d = {"cin": rb.in_channels, "cout": rb.out_channels,
                 "gn1": (f32(rb.norm1.weight), f32(rb.norm1.bias)), "gn2": (f32(rb.norm2.weight), f32(rb.norm2.bias)),
                 "conv1": pc(rb.conv1)}
            w2, b2 = pc(rb.conv2)
            d["skip"] = rb.in_channels != rb.out_channels
            if d["skip"]:
                ws, bs = pc(rb.nin_shortcut)
                w2, b2 = torch.cat([w2, ws], dim=1).contiguous(), b2 + bs
            d["conv2"] = (w2, b2)
            return d

        def attn(ab: AttnBlock):
            return {"c": ab.in_channels, "norm": (f32(ab.norm.weight), f32(ab.norm.bias)),
                    "q": pc(ab.q), "k": pc(ab.k), "v": pc(ab.v), "proj": pc(ab.proj_out)}

        def coder(net, is_enc):
            # bf16: input channels zero-padded to one 64-channel K block, so that conv_in runs on the tcgen05 kernel
            d = {"conv_in": pc(net.conv_in, pad_cin_to=self._cin_pad), "mid1": res(net.mid.block_1), "attn": attn(net.mid.attn_1),
                 "mid2": res(net.mid.block_2), "norm_out": (f32(net.norm_out.weight), f32(net.norm_out.bias)),
                 "conv_out": pc(net.conv_out), "levels": []}
            for lvl in (net.down if is_enc else net.up):
                e = {"blocks": [res(b) for b in lvl.block], "attn": [attn(a) for a in lvl.attn]}
                if hasattr(lvl, "downsample"):
                    e["resample"] = pc(lvl.downsample.conv)
                if hasattr(lvl, "upsample"):
                    e["resample"] = pc(lvl.upsample.conv)
                    if dtype == torch.bfloat16 and lvl.upsample.conv.in_channels % 64 == 0:
                        # nearest-2x + conv3x3 as four 2x2 output phases over the low-resolution input (4/9 of the FLOPs)
                        e["resample_phases"] = pack_upsample_phases(lvl.upsample.conv.weight, dtype)
                d["levels"].append(e)
            return d

        self.enc = coder(m.encoder, True)
        self.dec = coder(m.decoder, False)
        self.quant = pc(m.quant_conv)
        self.post_quant = pc(m.post_quant_conv, pad_cin_to=4)

    # Residual stream in fp32 with compute-dtype operand shadows, exactly as in unet.UNetEngine.
    def _new(self, n, h, w, c, dtype=None) -> Act:
        return Act.empty(n, h, w, c, dtype or self.dt, self.dev)

    def _new_dual(self, n, h, w, c) -> Dual:
        f = self._gp(self._new(n, h, w, c, torch.float32))
        return Dual(f, f if self.dt == torch.float32 else self._new(n, h, w, c))

    def _gp(self, f: Act) -> Act:
        """bf16 mode: the tcgen05 epilogue that writes this residual-stream tensor also writes its GroupNorm partial
        statistics (per 32-pixel chunk and 8-channel octet), so the GroupNorm reading it is one streaming pass instead
        of a statistics pass + an apply pass (as in unet.UNetEngine).  Octets resolve groups of >= 8 channels (256 / 512
        channel tensors); the 128-channel level (4 channels per group) gets one entry per channel quad."""
        if self.dt == torch.bfloat16 and f.c % 256 == 0:
            f.with_gn_partial()
        elif self.dt == torch.bfloat16 and f.c % 128 == 0:
            f.with_gn_partial(unit=4)        # 4 channels per group: one partial entry per channel quad
        return f

    @staticmethod
    def _out2(d: Dual):
        return None if d.h is d.f else d.h

    def _gn(self, x: Act, wb, silu: bool) -> Act:
        out = self._new(x.n, x.h, x.w, x.c)
        return ops.group_norm(x, wb[0], wb[1], 1e-6, out, self.stats, silu=silu)

    def _res(self, d, x: Dual) -> Dual:
        n, h, w = x.f.n, x.f.h, x.f.w
        h1 = self._gp(self._new(n, h, w, d["cout"], torch.float32))   # bf16 here costs the encoder its 1e-2 bound (measured 1.03e-2)
        ops.conv([ConvIn(self._gn(x.f, d["gn1"], True), 3, 1, 1)], d["conv1"][0], h1, bias=d["conv1"][1])
        hn2 = self._gn(h1, d["gn2"], True)
        out = self._new_dual(n, h, w, d["cout"])
        if d["skip"]:
            ops.conv([ConvIn(hn2, 3, 1, 1), ConvIn(x.h, 1, 1, 0)], d["conv2"][0], out.f, bias=d["conv2"][1],
                     out2=self._out2(out))
        else:
            ops.conv([ConvIn(hn2, 3, 1, 1)], d["conv2"][0], out.f, bias=d["conv2"][1], residual=x.f,
                     out2=self._out2(out))
        return out

    def _attn(self, d, x: Dual) -> Dual:
        """Single-head attention over all channels (d = C = 512): GEMM -> row softmax -> GEMM per image."""
        C_, n, hh, ww = d["c"], x.f.n, x.f.h, x.f.w
        tok = hh * ww
        xn = self._gn(x.f, d["norm"], False)
        q, k = self._new(n, hh, ww, C_), self._new(n, hh, ww, C_)
        ops.linear(xn, d["q"][0], q, bias=d["q"][1])
        ops.linear(xn, d["k"][0], k, bias=d["k"][1])
        o = self._new(n, hh, ww, C_)
        s = Act.empty(1, 1, tok, tok, self.dt, self.dev)
        vt = Act.empty(1, 1, C_, tok, self.dt, self.dev)
        wv = Act(d["v"][0], 1, 1, C_)
        for b in range(n):
            rows = slice(b * tok, (b + 1) * tok)
            qb = Act(q.buf[rows], 1, 1, tok)
            kb, xb = k.buf[rows], xn.buf[rows]
            ops.linear(qb, kb, s)                                  # S = q k^T
            ops.softmax_rows_(s, float(int(C_) ** (-0.5)))          # model.py:190-192
            ops.linear(wv, xb, vt)                                 # V^T = Wv xn^T (bias added below)
            ob = Act(o.buf[rows], 1, 1, tok)
            ops.linear(s, vt.buf, ob, bias=d["v"][1])              # P (V + 1 bv^T) = P V + bv (rows of P sum to 1)
        out = self._new_dual(n, hh, ww, C_)
        ops.linear(o, d["proj"][0], out.f, bias=d["proj"][1], residual=x.f, out2=self._out2(out))
        return out

    def _conv3(self, x: Act, wb, cout, stride=1, pad=1, upsample=False, w_phases=None) -> Dual:
        """3x3 conv of a compute-dtype operand into a fresh residual-stream tensor."""
        if stride == 2:
            ho, wo = x.h // 2, x.w // 2
        elif upsample:
            ho, wo = x.h * 2, x.w * 2
        else:
            ho, wo = x.h, x.w
        out = self._new_dual(x.n, ho, wo, cout)
        pow2 = lambda v: v & (v - 1) == 0  # noqa: E731
        if upsample and w_phases is not None and x.h * x.w >= 32 and pow2(x.h) and pow2(x.w):
            ops.conv([ConvIn(x, 3, 1, 1, upsample=1)], w_phases, out.f, bias=wb[1], out2=self._out2(out),
                     upsample_phases=True)
        elif upsample and self.dt == torch.bfloat16:
            up = self._new(x.n, ho, wo, x.c)
            ops.upsample_nearest2x(x, up)
            ops.conv([ConvIn(up, 3, 1, 1)], wb[0], out.f, bias=wb[1], out2=self._out2(out))
        else:
            ops.conv([ConvIn(x, 3, stride, pad, upsample=1 if upsample else 0)], wb[0], out.f, bias=wb[1],
                     out2=self._out2(out))
        return out

    def _input(self, x: torch.Tensor, cp: Optional[int] = None) -> Act:
        n, c, h, w = x.shape
        cp = cp or self._cin_pad
        buf = torch.zeros((n * h * w, cp), dtype=self.dt, device=self.dev)  # channels zero-padded (zero weights)
        a = Act(buf, n, h, w, c, 0)
        ops.nchw_to_nhwc(x.float().contiguous(), a)
        return Act(buf, n, h, w, cp, 0)

    def _output(self, h: Act, wb, cout) -> Act:
        co_pad = (cout + 7) // 8 * 8
        obuf = Act.empty(h.n, h.h, h.w, co_pad, torch.float32, self.dev)
        ops.conv([ConvIn(h, 3, 1, 1)], wb[0], obuf.cols(0, cout), bias=wb[1])
        return obuf.cols(0, cout)

    @torch.no_grad()
    def encode_moments(self, x: torch.Tensor) -> torch.Tensor:
        """Encoder.forward (model.py:434-459) + quant_conv (autoencoder.py:324-327) -> moments [N, 2*embed, h, w]."""
        hz = self.encoder_features(x)
        mo = self.quant[0].shape[0]
        mbuf = Act.empty(hz.n, hz.h, hz.w, (mo + 7) // 8 * 8, torch.float32, self.dev)
        ops.linear(hz, self.quant[0], mbuf.cols(0, mo), bias=self.quant[1])
        y = torch.empty((hz.n, mo, hz.h, hz.w), dtype=torch.float32, device=self.dev)
        return ops.nhwc_to_nchw(mbuf.cols(0, mo), y)

    @torch.no_grad()
    def encoder_features(self, x: torch.Tensor) -> Act:
        """Encoder.forward alone (model.py:434-459), NHWC in the compute dtype: what quant_conv reads, and what the EALDM
        conditioner takes from its `convs.encoder` (STDiff/models.py:515)."""
        E = self.enc
        n = x.shape[0]
        self.stats = ops.group_norm_workspace(n, x.shape[2] * x.shape[3], 4 * self.m.encoder.ch, self.dev)
        h = self._conv3(self._input(x), E["conv_in"], self.m.encoder.ch)
        for lvl in E["levels"]:
            for i, rb in enumerate(lvl["blocks"]):
                h = self._res(rb, h)
                if lvl["attn"]:
                    h = self._attn(lvl["attn"][i], h)
            if "resample" in lvl:
                h = self._conv3(h.h, lvl["resample"], h.f.c, stride=2, pad=0)
        h = self._res(E["mid1"], h)
        h = self._attn(E["attn"], h)
        h = self._res(E["mid2"], h)
        hn = self._gn(h.f, E["norm_out"], True)
        zc = E["conv_out"][0].shape[0]
        hz = self._new(hn.n, hn.h, hn.w, zc)
        ops.conv([ConvIn(hn, 3, 1, 1)], E["conv_out"][0], hz, bias=E["conv_out"][1])
        return hz

    @torch.no_grad()
    def decode(self, z: torch.Tensor) -> torch.Tensor:
        """post_quant_conv (autoencoder.py:330-333) + Decoder.forward (model.py:535-568)."""
        D = self.dec
        n = z.shape[0]
        up = 2 ** (self.m.decoder.num_resolutions - 1)
        self.stats = ops.group_norm_workspace(n, z.shape[2] * up * z.shape[3] * up, 4 * self.m.decoder.ch, self.dev)
        zin = self._input(z, 4)
        zc = self.post_quant[0].shape[0]
        cp = self._cin_pad
        zq = Act(torch.zeros((zin.rows, cp), dtype=self.dt, device=self.dev), zin.n, zin.h, zin.w, zc, 0)
        ops.linear(zin, self.post_quant[0], zq, bias=self.post_quant[1])
        h = self._conv3(Act(zq.buf, zq.n, zq.h, zq.w, cp, 0), D["conv_in"], D["conv_in"][0].shape[0])
        h = self._res(D["mid1"], h)
        h = self._attn(D["attn"], h)
        h = self._res(D["mid2"], h)
        for lvl in reversed(D["levels"]):
            for i, rb in enumerate(lvl["blocks"]):
                h = self._res(rb, h)
                if lvl["attn"]:
                    h = self._attn(lvl["attn"][i], h)
            if "resample" in lvl:
                h = self._conv3(h.h, lvl["resample"], h.f.c, upsample=True, w_phases=lvl.get("resample_phases"))
        hn = self._gn(h.f, D["norm_out"], True)
        co = D["conv_out"][0].shape[0]
        o = self._output(hn, D["conv_out"], co)
        y = torch.empty((o.n, co, o.h, o.w), dtype=torch.float32, device=self.dev)
        return ops.nhwc_to_nchw(o, y)
