"""Oracle: AutoencoderKL encoder / decoder (TEST INFRASTRUCTURE, see oracle/__init__.py).

Functional fp32 restatement over a flat state_dict with the reference's names
(`encoder.down.0.block.0.norm1.weight`, `decoder.up.3.upsample.conv.weight`, `quant_conv.weight`...)."""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


def _gn(sd: SD, p: str, x):  # Normalize: GroupNorm(32, eps=1e-6), model.py:38-39
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], 1e-6)


def resnet_block(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """ResnetBlock.forward with temb=None, model.py:116-141 (swish = x*sigmoid(x), dropout 0)."""
    h = F.conv2d(F.silu(_gn(sd, p + "norm1", x)), sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=1)
    h = F.conv2d(F.silu(_gn(sd, p + "norm2", h)), sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1)
    if p + "nin_shortcut.weight" in sd:
        x = F.conv2d(x, sd[p + "nin_shortcut.weight"], sd[p + "nin_shortcut.bias"])
    return x + h


def attn_block(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """AttnBlock.forward, model.py:175-202: single head over all channels, scale c^-0.5."""
    b, c, h, w = x.shape
    n = _gn(sd, p + "norm", x)
    q = F.conv2d(n, sd[p + "q.weight"], sd[p + "q.bias"]).reshape(b, c, h * w).permute(0, 2, 1)
    k = F.conv2d(n, sd[p + "k.weight"], sd[p + "k.bias"]).reshape(b, c, h * w)
    v = F.conv2d(n, sd[p + "v.weight"], sd[p + "v.bias"]).reshape(b, c, h * w)
    w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    o = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, h, w)
    return x + F.conv2d(o, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])


def encoder_forward(sd: SD, dd: dict, x: torch.Tensor, p: str = "encoder.") -> torch.Tensor:
    """Encoder.forward, model.py:434-459.  Downsample = F.pad(0,1,0,1) + conv s2 p0, model.py:72-76."""
    nres = len(dd["ch_mult"])
    res = dd["resolution"]
    h = F.conv2d(x, sd[p + "conv_in.weight"], sd[p + "conv_in.bias"], padding=1)
    for lvl in range(nres):
        for blk in range(dd["num_res_blocks"]):
            h = resnet_block(sd, f"{p}down.{lvl}.block.{blk}.", h)
            if res in dd["attn_resolutions"]:
                h = attn_block(sd, f"{p}down.{lvl}.attn.{blk}.", h)
        if lvl != nres - 1:
            h = F.pad(h, (0, 1, 0, 1), mode="constant", value=0)
            h = F.conv2d(h, sd[f"{p}down.{lvl}.downsample.conv.weight"], sd[f"{p}down.{lvl}.downsample.conv.bias"],
                         stride=2)
            res //= 2
    h = resnet_block(sd, p + "mid.block_1.", h)
    h = attn_block(sd, p + "mid.attn_1.", h)
    h = resnet_block(sd, p + "mid.block_2.", h)
    h = F.silu(_gn(sd, p + "norm_out", h))
    return F.conv2d(h, sd[p + "conv_out.weight"], sd[p + "conv_out.bias"], padding=1)


def decoder_forward(sd: SD, dd: dict, z: torch.Tensor, p: str = "decoder.") -> torch.Tensor:
    """Decoder.forward, model.py:535-568."""
    nres = len(dd["ch_mult"])
    res = dd["resolution"] // 2 ** (nres - 1)
    h = F.conv2d(z, sd[p + "conv_in.weight"], sd[p + "conv_in.bias"], padding=1)
    h = resnet_block(sd, p + "mid.block_1.", h)
    h = attn_block(sd, p + "mid.attn_1.", h)
    h = resnet_block(sd, p + "mid.block_2.", h)
    for lvl in reversed(range(nres)):
        for blk in range(dd["num_res_blocks"] + 1):
            h = resnet_block(sd, f"{p}up.{lvl}.block.{blk}.", h)
            if res in dd["attn_resolutions"]:
                h = attn_block(sd, f"{p}up.{lvl}.attn.{blk}.", h)
        if lvl != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = F.conv2d(h, sd[f"{p}up.{lvl}.upsample.conv.weight"], sd[f"{p}up.{lvl}.upsample.conv.bias"], padding=1)
            res *= 2
    h = F.silu(_gn(sd, p + "norm_out", h))
    return F.conv2d(h, sd[p + "conv_out.weight"], sd[p + "conv_out.bias"], padding=1)


def kl_encode_moments(sd: SD, dd: dict, x: torch.Tensor) -> torch.Tensor:
    """AutoencoderKL.encode up to the moments, autoencoder.py:324-327 (quant_conv 1x1)."""
    return F.conv2d(encoder_forward(sd, dd, x), sd["quant_conv.weight"], sd["quant_conv.bias"])


def gaussian_from_moments(moments: torch.Tensor):
    """DiagonalGaussianDistribution.__init__, distributions.py:24-33: (mean, logvar clamped to [-30,20], std)."""
    mean, logvar = torch.chunk(moments, 2, dim=1)
    logvar = torch.clamp(logvar, -30.0, 20.0)
    return mean, logvar, torch.exp(0.5 * logvar)


def kl_decode(sd: SD, dd: dict, z: torch.Tensor) -> torch.Tensor:
    """AutoencoderKL.decode, autoencoder.py:330-333 (post_quant_conv 1x1 then Decoder)."""
    return decoder_forward(sd, dd, F.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"]))


def decode_first_stage(sd: SD, dd: dict, z: torch.Tensor, scale_factor: float = 1.0) -> torch.Tensor:
    """LatentDiffusion.decode_first_stage, ddpm.py:721,768-771: z / scale_factor -> first stage decode."""
    return kl_decode(sd, dd, 1.0 / scale_factor * z)


def vq_decode(sd: SD, dd: dict, h: torch.Tensor, force_not_quantize: bool = False) -> torch.Tensor:
    """VQModelInterface.decode, autoencoder.py:274-282: quantize (oracle/vq.py) -> post_quant_conv -> Decoder."""
    from .vq import vq_nearest
    quant = h if force_not_quantize else vq_nearest(h, sd["quantize.embedding.weight"])[0]
    return decoder_forward(sd, dd, F.conv2d(quant, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"]))


def vq_encode(sd: SD, dd: dict, x: torch.Tensor) -> torch.Tensor:
    """VQModelInterface.encode, autoencoder.py:268-271: encoder -> quant_conv (no quantisation)."""
    return F.conv2d(encoder_forward(sd, dd, x), sd["quant_conv.weight"], sd["quant_conv.bias"])


def vq_param_shapes(dd: dict, embed_dim: int, n_embed: int) -> List[Tuple[str, Tuple[int, ...]]]:
    """VQModel registration order (autoencoder.py:35-43): encoder, decoder, quantize, quant_conv, post_quant_conv."""
    kl = autoencoder_kl_param_shapes(dict(dd, double_z=False), embed_dim)
    body = [e for e in kl if not e[0].startswith(("quant_conv", "post_quant_conv"))]
    zc = dd["z_channels"]
    return body + [("quantize.embedding.weight", (n_embed, embed_dim)),
                   ("quant_conv.weight", (embed_dim, zc, 1, 1)), ("quant_conv.bias", (embed_dim,)),
                   ("post_quant_conv.weight", (zc, embed_dim, 1, 1)), ("post_quant_conv.bias", (zc,))]


# ---- parameter inventory (registration order of the reference constructors) ---------------------------
def _res_shapes(p, cin, cout):
    s = [(p + "norm1.weight", (cin,)), (p + "norm1.bias", (cin,)),
         (p + "conv1.weight", (cout, cin, 3, 3)), (p + "conv1.bias", (cout,)),
         (p + "norm2.weight", (cout,)), (p + "norm2.bias", (cout,)),
         (p + "conv2.weight", (cout, cout, 3, 3)), (p + "conv2.bias", (cout,))]
    if cin != cout:
        s += [(p + "nin_shortcut.weight", (cout, cin, 1, 1)), (p + "nin_shortcut.bias", (cout,))]
    return s


def _attn_shapes(p, c):
    s = [(p + "norm.weight", (c,)), (p + "norm.bias", (c,))]
    for n in ("q", "k", "v", "proj_out"):
        s += [(f"{p}{n}.weight", (c, c, 1, 1)), (f"{p}{n}.bias", (c,))]
    return s


def autoencoder_kl_param_shapes(dd: dict, embed_dim: int) -> List[Tuple[str, Tuple[int, ...]]]:
    ch, mult, nrb = dd["ch"], list(dd["ch_mult"]), dd["num_res_blocks"]
    zc = dd["z_channels"]
    nres = len(mult)
    out: List[Tuple[str, Tuple[int, ...]]] = []
    # Encoder (model.py:368-432)
    p = "encoder."
    out += [(p + "conv_in.weight", (ch, dd["in_channels"], 3, 3)), (p + "conv_in.bias", (ch,))]
    in_mult = [1] + mult
    res = dd["resolution"]
    block_in = ch
    for lvl in range(nres):
        block_in = ch * in_mult[lvl]
        block_out = ch * mult[lvl]
        for blk in range(nrb):
            out += _res_shapes(f"{p}down.{lvl}.block.{blk}.", block_in, block_out)
            block_in = block_out
        if res in dd["attn_resolutions"]:
            for blk in range(nrb):
                out += _attn_shapes(f"{p}down.{lvl}.attn.{blk}.", block_in)
        if lvl != nres - 1:
            out += [(f"{p}down.{lvl}.downsample.conv.weight", (block_in, block_in, 3, 3)),
                    (f"{p}down.{lvl}.downsample.conv.bias", (block_in,))]
            res //= 2
    out += _res_shapes(p + "mid.block_1.", block_in, block_in)
    out += _attn_shapes(p + "mid.attn_1.", block_in)
    out += _res_shapes(p + "mid.block_2.", block_in, block_in)
    out += [(p + "norm_out.weight", (block_in,)), (p + "norm_out.bias", (block_in,))]
    zo = 2 * zc if dd.get("double_z", True) else zc
    out += [(p + "conv_out.weight", (zo, block_in, 3, 3)), (p + "conv_out.bias", (zo,))]
    # Decoder (model.py:462-533); `up` is built from the lowest resolution and prepended
    p = "decoder."
    block_in = ch * mult[nres - 1]
    res = dd["resolution"] // 2 ** (nres - 1)
    out += [(p + "conv_in.weight", (block_in, zc, 3, 3)), (p + "conv_in.bias", (block_in,))]
    out += _res_shapes(p + "mid.block_1.", block_in, block_in)
    out += _attn_shapes(p + "mid.attn_1.", block_in)
    out += _res_shapes(p + "mid.block_2.", block_in, block_in)
    ups = {}
    for lvl in reversed(range(nres)):
        cur = []
        block_out = ch * mult[lvl]
        for blk in range(nrb + 1):
            cur += _res_shapes(f"{p}up.{lvl}.block.{blk}.", block_in, block_out)
            block_in = block_out
        if res in dd["attn_resolutions"]:
            for blk in range(nrb + 1):
                cur += _attn_shapes(f"{p}up.{lvl}.attn.{blk}.", block_in)
        if lvl != 0:
            cur += [(f"{p}up.{lvl}.upsample.conv.weight", (block_in, block_in, 3, 3)),
                    (f"{p}up.{lvl}.upsample.conv.bias", (block_in,))]
            res *= 2
        ups[lvl] = cur
    for lvl in range(nres):
        out += ups[lvl]
    out += [(p + "norm_out.weight", (block_in,)), (p + "norm_out.bias", (block_in,))]
    out += [(p + "conv_out.weight", (dd["out_ch"], block_in, 3, 3)), (p + "conv_out.bias", (dd["out_ch"],))]
    # AutoencoderKL (autoencoder.py:302-303)
    out += [("quant_conv.weight", (2 * embed_dim, 2 * zc, 1, 1)), ("quant_conv.bias", (2 * embed_dim,)),
            ("post_quant_conv.weight", (zc, embed_dim, 1, 1)), ("post_quant_conv.bias", (zc,))]
    return out
