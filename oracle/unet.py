"""Oracle: functional fp32 restatement of the reference UNet (TEST INFRASTRUCTURE, see oracle/__init__.py).

All functions take a flat state_dict `sd` with the reference's parameter names
(`input_blocks.1.0.in_layers.2.weight`, ...) and NCHW tensors, exactly like the reference modules.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# ---- ldm/modules/diffusionmodules/util.py:151-171 -------------------------------------------------
def timestep_embedding(timesteps: torch.Tensor, dim: int, max_period: float = 10000.0) -> torch.Tensor:
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = timesteps[:, None].float() * freqs[None].to(timesteps.device)
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)  # cos first, then sin
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


# ---- structure: mirrors UNetModel.__init__, openaimodel.py:506-692 ---------------------------------
def unet_structure(cfg: dict) -> dict:
    """Returns {'input': [[layer,...],...], 'middle': [...], 'output': [[...],...]} where a layer is
    ('conv_in', cin, cout) | ('res', cin, cout) | ('attn', ch, heads, dim_head) | ('down', ch) | ('up', ch)."""
    mc = cfg["model_channels"]
    mult = list(cfg.get("channel_mult", (1, 2, 4, 8)))
    nrb = cfg["num_res_blocks"]
    attn_res = set(cfg["attention_resolutions"])
    num_heads = cfg.get("num_heads", -1)
    nhc = cfg.get("num_head_channels", -1)
    use_st = cfg.get("use_spatial_transformer", False)
    legacy = cfg.get("legacy", True)

    def heads_for(ch):  # openaimodel.py:541-549 (and the copies at :589-596, :642-650)
        nonlocal num_heads
        if nhc == -1:
            dim_head = ch // num_heads
        else:
            num_heads = ch // nhc
            dim_head = nhc
        if legacy:
            dim_head = ch // num_heads if use_st else nhc
        if not use_st:
            # AttentionBlock(num_heads=num_heads, num_head_channels=dim_head), openaimodel.py:295-301
            h = num_heads if dim_head == -1 else ch // dim_head
            return h, ch // h
        return num_heads, dim_head

    inp: List[list] = [[("conv_in", cfg["in_channels"], mc)]]
    chans = [mc]
    ch, ds = mc, 1
    for level, m in enumerate(mult):
        for _ in range(nrb):
            layers = [("res", ch, m * mc)]
            ch = m * mc
            if ds in attn_res:
                layers.append(("attn", ch) + heads_for(ch))
            inp.append(layers)
            chans.append(ch)
        if level != len(mult) - 1:
            inp.append([("down", ch)])
            chans.append(ch)
            ds *= 2
    mid = [("res", ch, ch), ("attn", ch) + heads_for(ch), ("res", ch, ch)]
    out: List[list] = []
    for level, m in list(enumerate(mult))[::-1]:
        for i in range(nrb + 1):
            ich = chans.pop()
            layers = [("res", ch + ich, mc * m)]
            ch = mc * m
            if ds in attn_res:
                layers.append(("attn", ch) + heads_for(ch))
            if level and i == nrb:
                layers.append(("up", ch))
                ds //= 2
            out.append(layers)
    return {"input": inp, "middle": mid, "output": out, "out_ch": ch}


# ---- blocks -----------------------------------------------------------------------------------------
def res_block(sd: SD, p: str, x: torch.Tensor, emb: torch.Tensor) -> torch.Tensor:
    """ResBlock._forward, openaimodel.py:255-275 (no up/down, no scale-shift norm). GroupNorm32 eps 1e-5."""
    h = F.group_norm(x.float(), 32, sd[p + "in_layers.0.weight"], sd[p + "in_layers.0.bias"], 1e-5)
    h = F.conv2d(F.silu(h), sd[p + "in_layers.2.weight"], sd[p + "in_layers.2.bias"], padding=1)
    e = F.linear(F.silu(emb), sd[p + "emb_layers.1.weight"], sd[p + "emb_layers.1.bias"])
    h = h + e[:, :, None, None]
    h = F.group_norm(h.float(), 32, sd[p + "out_layers.0.weight"], sd[p + "out_layers.0.bias"], 1e-5)
    h = F.conv2d(F.silu(h), sd[p + "out_layers.3.weight"], sd[p + "out_layers.3.bias"], padding=1)
    if p + "skip_connection.weight" in sd:
        w = sd[p + "skip_connection.weight"]
        x = F.conv2d(x, w, sd[p + "skip_connection.bias"], padding=w.shape[-1] // 2)
    return x + h


def attention_block(sd: SD, p: str, x: torch.Tensor, heads: int) -> torch.Tensor:
    """AttentionBlock._forward + QKVAttentionLegacy.forward, openaimodel.py:318-324, 356-372."""
    b, c, hh, ww = x.shape
    xf = x.reshape(b, c, -1)
    n = F.group_norm(xf.float(), 32, sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)
    qkv = F.conv1d(n, sd[p + "qkv.weight"], sd[p + "qkv.bias"])  # [b, 3c, T]
    ch = c // heads
    q, k, v = qkv.reshape(b * heads, ch * 3, -1).split(ch, dim=1)  # per-head [q;k;v] interleave
    scale = 1 / math.sqrt(math.sqrt(ch))
    w = torch.einsum("bct,bcs->bts", q * scale, k * scale)
    w = torch.softmax(w.float(), dim=-1)
    a = torch.einsum("bts,bcs->bct", w, v).reshape(b, -1, xf.shape[-1])
    h = F.conv1d(a, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])
    return (xf + h).reshape(b, c, hh, ww)


def cross_attention(sd: SD, p: str, x: torch.Tensor, context: Optional[torch.Tensor], heads: int) -> torch.Tensor:
    """CrossAttention.forward, attention.py:170-193 (mask=None). x: [b, n, c]."""
    ctx = x if context is None else context
    q = F.linear(x, sd[p + "to_q.weight"])
    k = F.linear(ctx, sd[p + "to_k.weight"])
    v = F.linear(ctx, sd[p + "to_v.weight"])
    b, n, inner = q.shape
    d = inner // heads

    def split(t):
        return t.reshape(b, t.shape[1], heads, d).permute(0, 2, 1, 3)

    q, k, v = split(q), split(k), split(v)
    sim = torch.einsum("bhid,bhjd->bhij", q, k) * (d ** -0.5)
    out = torch.einsum("bhij,bhjd->bhid", sim.softmax(dim=-1), v)
    out = out.permute(0, 2, 1, 3).reshape(b, n, inner)
    return F.linear(out, sd[p + "to_out.0.weight"], sd[p + "to_out.0.bias"])


def feed_forward(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """FeedForward with GEGLU, attention.py:37-64: Linear(c->8c), value*gelu_erf(gate), Linear(4c->c)."""
    y = F.linear(x, sd[p + "net.0.proj.weight"], sd[p + "net.0.proj.bias"])
    val, gate = y.chunk(2, dim=-1)
    return F.linear(val * F.gelu(gate), sd[p + "net.2.weight"], sd[p + "net.2.bias"])


def basic_transformer_block(sd: SD, p: str, x: torch.Tensor, context, heads: int) -> torch.Tensor:
    """BasicTransformerBlock._forward, attention.py:211-215 (LayerNorm eps 1e-5)."""
    c = x.shape[-1]

    def ln(t, name):
        return F.layer_norm(t, (c,), sd[p + name + ".weight"], sd[p + name + ".bias"], 1e-5)

    x = cross_attention(sd, p + "attn1.", ln(x, "norm1"), None, heads) + x
    x = cross_attention(sd, p + "attn2.", ln(x, "norm2"), context, heads) + x
    x = feed_forward(sd, p + "ff.", ln(x, "norm3")) + x
    return x


def spatial_transformer(sd: SD, p: str, x: torch.Tensor, context, heads: int, depth: int = 1) -> torch.Tensor:
    """SpatialTransformer.forward, attention.py:250-261 (GroupNorm eps 1e-6, attention.py:76-77)."""
    b, c, h, w = x.shape
    y = F.group_norm(x, 32, sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-6)
    y = F.conv2d(y, sd[p + "proj_in.weight"], sd[p + "proj_in.bias"])
    y = y.permute(0, 2, 3, 1).reshape(b, h * w, -1)
    for i in range(depth):
        y = basic_transformer_block(sd, f"{p}transformer_blocks.{i}.", y, context, heads)
    y = y.reshape(b, h, w, -1).permute(0, 3, 1, 2)
    y = F.conv2d(y, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])
    return y + x


def _run_layers(sd: SD, prefix: str, layers: Sequence[tuple], h, emb, context, use_st: bool, depth: int):
    for j, layer in enumerate(layers):
        p = f"{prefix}{j}."
        kind = layer[0]
        if kind == "conv_in":
            h = F.conv2d(h, sd[p + "weight"], sd[p + "bias"], padding=1)
        elif kind == "res":
            h = res_block(sd, p, h, emb)
        elif kind == "attn":
            heads = layer[2]
            h = spatial_transformer(sd, p, h, context, heads, depth) if use_st else attention_block(sd, p, h, heads)
        elif kind == "down":  # Downsample: conv3x3 stride 2 pad 1, openaimodel.py:149-160
            h = F.conv2d(h, sd[p + "op.weight"], sd[p + "op.bias"], stride=2, padding=1)
        elif kind == "up":  # Upsample: nearest 2x then conv3x3 pad 1, openaimodel.py:109-119
            h = F.interpolate(h, scale_factor=2, mode="nearest")
            h = F.conv2d(h, sd[p + "conv.weight"], sd[p + "conv.bias"], padding=1)
        else:
            raise ValueError(kind)
    return h


def unet_forward(sd: SD, cfg: dict, x: torch.Tensor, timesteps: torch.Tensor,
                 context: Optional[torch.Tensor] = None) -> torch.Tensor:
    """UNetModel.forward, openaimodel.py:710-742."""
    st = unet_structure(cfg)
    use_st = cfg.get("use_spatial_transformer", False)
    depth = cfg.get("transformer_depth", 1)
    t_emb = timestep_embedding(timesteps, cfg["model_channels"])
    emb = F.linear(t_emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    hs = []
    h = x.float()
    for i, layers in enumerate(st["input"]):
        h = _run_layers(sd, f"input_blocks.{i}.", layers, h, emb, context, use_st, depth)
        hs.append(h)
    h = _run_layers(sd, "middle_block.", st["middle"], h, emb, context, use_st, depth)
    for i, layers in enumerate(st["output"]):
        h = torch.cat([h, hs.pop()], dim=1)
        h = _run_layers(sd, f"output_blocks.{i}.", layers, h, emb, context, use_st, depth)
    h = F.group_norm(h.float(), 32, sd["out.0.weight"], sd["out.0.bias"], 1e-5)
    return F.conv2d(F.silu(h), sd["out.2.weight"], sd["out.2.bias"], padding=1)


# ---- deterministic synthetic weights ---------------------------------------------------------------
def unet_param_shapes(cfg: dict) -> List[Tuple[str, Tuple[int, ...]]]:
    """Names and shapes of the reference UNet's parameters in the reference's registration order
    (checked against the reference constructor by tests/test_oracle_golden.py)."""
    st = unet_structure(cfg)
    mc = cfg["model_channels"]
    ted = mc * 4
    use_st = cfg.get("use_spatial_transformer", False)
    ctx = cfg.get("context_dim", None)
    depth = cfg.get("transformer_depth", 1)
    out: List[Tuple[str, Tuple[int, ...]]] = []

    def wb(name, wshape):
        out.append((name + ".weight", tuple(wshape)))
        out.append((name + ".bias", (wshape[0],)))

    wb("time_embed.0", (ted, mc))
    wb("time_embed.2", (ted, ted))

    def emit(prefix, layers):
        for j, layer in enumerate(layers):
            p = f"{prefix}{j}."
            kind = layer[0]
            if kind == "conv_in":
                wb(p[:-1], (layer[2], layer[1], 3, 3))
            elif kind == "res":
                cin, cout = layer[1], layer[2]
                wb(p + "in_layers.0", (cin,))
                wb(p + "in_layers.2", (cout, cin, 3, 3))
                wb(p + "emb_layers.1", (cout, ted))
                wb(p + "out_layers.0", (cout,))
                wb(p + "out_layers.3", (cout, cout, 3, 3))
                if cin != cout:
                    wb(p + "skip_connection", (cout, cin, 1, 1))
            elif kind == "attn":
                ch, heads, dh = layer[1], layer[2], layer[3]
                if use_st:
                    inner = heads * dh
                    wb(p + "norm", (ch,))
                    wb(p + "proj_in", (inner, ch, 1, 1))
                    for d in range(depth):
                        q = f"{p}transformer_blocks.{d}."
                        out.append((q + "attn1.to_q.weight", (inner, inner)))
                        out.append((q + "attn1.to_k.weight", (inner, inner)))
                        out.append((q + "attn1.to_v.weight", (inner, inner)))
                        wb(q + "attn1.to_out.0", (inner, inner))
                        wb(q + "ff.net.0.proj", (inner * 8, inner))
                        wb(q + "ff.net.2", (inner, inner * 4))
                        out.append((q + "attn2.to_q.weight", (inner, inner)))
                        out.append((q + "attn2.to_k.weight", (inner, ctx)))
                        out.append((q + "attn2.to_v.weight", (inner, ctx)))
                        wb(q + "attn2.to_out.0", (inner, inner))
                        wb(q + "norm1", (inner,))
                        wb(q + "norm2", (inner,))
                        wb(q + "norm3", (inner,))
                    wb(p + "proj_out", (ch, inner, 1, 1))
                else:
                    wb(p + "norm", (ch,))
                    wb(p + "qkv", (3 * ch, ch, 1))
                    wb(p + "proj_out", (ch, ch, 1))
            elif kind == "down":
                wb(p + "op", (layer[1], layer[1], 3, 3))
            elif kind == "up":
                wb(p + "conv", (layer[1], layer[1], 3, 3))

    for i, layers in enumerate(st["input"]):
        emit(f"input_blocks.{i}.", layers)
    emit("middle_block.", st["middle"])
    for i, layers in enumerate(st["output"]):
        emit(f"output_blocks.{i}.", layers)
    wb("out.0", (st["out_ch"],))
    wb("out.2", (cfg["out_channels"], mc, 3, 3))
    return out


def synthetic_state_dict(shapes: Sequence[Tuple[str, Tuple[int, ...]]], seed: int) -> SD:
    """Deterministic, reference-independent 'random-init' weights with the DISTRIBUTION BASELINE.md
    section 4 specifies: the reference constructors' initialisation (nn.Conv2d / nn.Linear default =
    kaiming_uniform(a=sqrt 5), i.e. U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weights and biases; norm
    gains 1, norm biases 0) followed by N(0, 0.02) on the tensors the reference zero-initialises
    (ResBlock.out_layers.3, *.proj_out, out.2 -- openaimodel.py:229-231,312,685; attention.py:244-248),
    whose biases stay 0.  Without that re-randomisation the UNet output is identically 0 and every
    relative metric is 0/0.  Each tensor has its own torch.Generator stream
    (seed, index), so the values do not depend on the reference being importable."""
    by_name = dict(shapes)

    def zero_init(wname: str) -> bool:
        return wname.endswith("out_layers.3.weight") or wname.endswith("proj_out.weight") or wname == "out.2.weight"

    sd: SD = {}
    for idx, (name, shape) in enumerate(shapes):
        gen = torch.Generator().manual_seed(seed * 100003 + idx)
        if name.endswith(".bias"):
            wshape = by_name.get(name[:-5] + ".weight")
            if wshape is None or len(wshape) == 1 or zero_init(name[:-5] + ".weight"):
                t = torch.zeros(shape)
            else:
                bound = 1.0 / math.sqrt(math.prod(wshape[1:]))
                t = (torch.rand(shape, generator=gen) * 2 - 1) * bound
        elif len(shape) == 1:  # norm gain
            t = torch.ones(shape)
        elif zero_init(name):
            t = 0.02 * torch.randn(shape, generator=gen)
        else:
            bound = 1.0 / math.sqrt(math.prod(shape[1:]))
            t = (torch.rand(shape, generator=gen) * 2 - 1) * bound
        sd[name] = t
    return sd
