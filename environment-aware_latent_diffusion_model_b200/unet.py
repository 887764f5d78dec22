"""Drop-in `UNetModel` for EALDM's `unet_config.target` (reference:
ldm/modules/diffusionmodules/openaimodel.py:413-742).

The module tree below exists to own the parameters under the reference's names and shapes (so
reference checkpoints load with strict=True and EMA / optimizers see the same tensors); the leaf
`nn.Conv2d` / `nn.Linear` / `nn.GroupNorm` / `nn.LayerNorm` objects are never called.  `forward`
hands the whole network to `UNetEngine`, which packs the weights to kernel layout once and then
issues libealdm_b200 kernels on NHWC activations (see DESIGN.md for the data layout).
"""
from __future__ import annotations

import contextlib
import math
import os
from typing import List, Optional

import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .ops import Act, ConvIn
from .packing import collapse_cross_attention, geglu_interleave, pack_conv_weight, pack_upsample_phases


def zero_module(module: nn.Module) -> nn.Module:
    """Reference initialisation of output convs (openaimodel.py:229-231,312,685; attention.py:244)."""
    for p in module.parameters():
        p.detach().zero_()
    return module


# ---- parameter-owning modules (names mirror the reference) -----------------------------------------
class TimestepEmbedSequential(nn.Sequential):
    """openaimodel.py:74-88"""


class Upsample(nn.Module):
    """openaimodel.py:91-119: nearest 2x then conv3x3."""

    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1):
        super().__init__()
        self.channels, self.out_channels, self.use_conv = channels, out_channels or channels, use_conv
        if not use_conv:
            raise NotImplementedError("Upsample without convolution (conv_resample=False)")
        self.conv = nn.Conv2d(channels, self.out_channels, 3, padding=padding)


class Downsample(nn.Module):
    """openaimodel.py:134-160: conv3x3 stride 2 pad 1."""

    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1):
        super().__init__()
        self.channels, self.out_channels, self.use_conv = channels, out_channels or channels, use_conv
        if not use_conv:
            raise NotImplementedError("Downsample without convolution (conv_resample=False)")
        self.op = nn.Conv2d(channels, self.out_channels, 3, stride=2, padding=padding)


class ResBlock(nn.Module):
    """openaimodel.py:163-275"""

    def __init__(self, channels, emb_channels, dropout, out_channels=None):
        super().__init__()
        self.channels, self.out_channels = channels, out_channels or channels
        self.in_layers = nn.Sequential(nn.GroupNorm(32, channels), nn.SiLU(),
                                       nn.Conv2d(channels, self.out_channels, 3, padding=1))
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(emb_channels, self.out_channels))
        self.out_layers = nn.Sequential(nn.GroupNorm(32, self.out_channels), nn.SiLU(), nn.Dropout(p=dropout),
                                        zero_module(nn.Conv2d(self.out_channels, self.out_channels, 3, padding=1)))
        self.skip_connection = (nn.Identity() if self.out_channels == channels
                                else nn.Conv2d(channels, self.out_channels, 1))


class AttentionBlock(nn.Module):
    """openaimodel.py:278-324 with QKVAttentionLegacy (:347-372)."""

    def __init__(self, channels, num_heads=1, num_head_channels=-1):
        super().__init__()
        self.channels = channels
        self.num_heads = num_heads if num_head_channels == -1 else channels // num_head_channels
        self.norm = nn.GroupNorm(32, channels)
        self.qkv = nn.Conv1d(channels, channels * 3, 1)
        self.proj_out = zero_module(nn.Conv1d(channels, channels, 1))


class CrossAttention(nn.Module):
    """attention.py:152-193"""

    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64):
        super().__init__()
        inner = dim_head * heads
        context_dim = query_dim if context_dim is None else context_dim
        self.heads, self.dim_head, self.scale = heads, dim_head, dim_head ** -0.5
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(context_dim, inner, bias=False)
        self.to_v = nn.Linear(context_dim, inner, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, query_dim), nn.Dropout(0.0))


class GEGLU(nn.Module):
    """attention.py:37-44"""

    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)


class FeedForward(nn.Module):
    """attention.py:47-64 (glu=True)"""

    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = nn.Sequential(GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim))


class BasicTransformerBlock(nn.Module):
    """attention.py:196-215"""

    def __init__(self, dim, n_heads, d_head, context_dim=None):
        super().__init__()
        self.attn1 = CrossAttention(dim, heads=n_heads, dim_head=d_head)
        self.ff = FeedForward(dim)
        self.attn2 = CrossAttention(dim, context_dim=context_dim, heads=n_heads, dim_head=d_head)
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(dim), nn.LayerNorm(dim), nn.LayerNorm(dim)


class SpatialTransformer(nn.Module):
    """attention.py:218-261"""

    def __init__(self, in_channels, n_heads, d_head, depth=1, context_dim=None):
        super().__init__()
        self.in_channels, self.n_heads, self.d_head = in_channels, n_heads, d_head
        inner = n_heads * d_head
        self.norm = nn.GroupNorm(32, in_channels, eps=1e-6)
        self.proj_in = nn.Conv2d(in_channels, inner, 1)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, n_heads, d_head, context_dim=context_dim) for _ in range(depth)])
        self.proj_out = zero_module(nn.Conv2d(inner, in_channels, 1))


class UNetModel(nn.Module):
    """Same constructor contract as the reference UNetModel (openaimodel.py:443-469); combinations the
    EALDM configs never use raise NotImplementedError instead of silently diverging.
    Extra keyword (not in the reference): `compute_dtype` in {"bf16", "fp32"} selects the tensor-core
    path (default) or the fp32 parity path; it can also be switched later with `set_compute_dtype`."""

    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks,
                 attention_resolutions, dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2,
                 num_classes=None, use_checkpoint=False, use_fp16=False, num_heads=-1, num_head_channels=-1,
                 num_heads_upsample=-1, use_scale_shift_norm=False, resblock_updown=False,
                 use_new_attention_order=False, use_spatial_transformer=False, transformer_depth=1,
                 context_dim=None, n_embed=None, legacy=True, compute_dtype="bf16"):
        super().__init__()
        if use_spatial_transformer:
            assert context_dim is not None, "context_dim is required with use_spatial_transformer"
        if context_dim is not None:
            assert use_spatial_transformer, "context_dim requires use_spatial_transformer"
            if not isinstance(context_dim, int):
                context_dim = list(context_dim)
                assert len(context_dim) == 1, "one context_dim per transformer depth is not supported"
                context_dim = context_dim[0]
        for flag, name in ((dims != 2, "dims != 2"), (num_classes is not None, "num_classes"),
                           (use_scale_shift_norm, "use_scale_shift_norm"), (resblock_updown, "resblock_updown"),
                           (use_new_attention_order, "use_new_attention_order"), (n_embed is not None, "n_embed"),
                           (not conv_resample, "conv_resample=False"), (dropout != 0, "dropout != 0")):
            if flag:
                raise NotImplementedError(f"UNetModel option not supported by the B200 path: {name}")
        if num_heads_upsample == -1:
            num_heads_upsample = num_heads
        if num_heads == -1:
            assert num_head_channels != -1, "Either num_heads or num_head_channels has to be set"
        if num_head_channels == -1:
            assert num_heads != -1, "Either num_heads or num_head_channels has to be set"

        self.image_size, self.in_channels, self.model_channels = image_size, in_channels, model_channels
        self.out_channels, self.num_res_blocks = out_channels, num_res_blocks
        self.attention_resolutions, self.dropout = attention_resolutions, dropout
        self.channel_mult, self.conv_resample, self.num_classes = channel_mult, conv_resample, num_classes
        self.use_checkpoint = use_checkpoint
        self.dtype = torch.float16 if use_fp16 else torch.float32  # reference attribute (openaimodel.py:500)
        self.num_heads, self.num_head_channels, self.num_heads_upsample = num_heads, num_head_channels, num_heads_upsample
        self.predict_codebook_ids = False
        self.use_spatial_transformer, self.context_dim = use_spatial_transformer, context_dim

        ted = model_channels * 4
        self.time_embed = nn.Sequential(nn.Linear(model_channels, ted), nn.SiLU(), nn.Linear(ted, ted))

        def attn_layer(ch):
            nonlocal num_heads
            if num_head_channels == -1:
                dim_head = ch // num_heads
            else:
                num_heads = ch // num_head_channels
                dim_head = num_head_channels
            if legacy:
                dim_head = ch // num_heads if use_spatial_transformer else num_head_channels
            if use_spatial_transformer:
                return SpatialTransformer(ch, num_heads, dim_head, depth=transformer_depth, context_dim=context_dim)
            return AttentionBlock(ch, num_heads=num_heads, num_head_channels=dim_head)

        self.input_blocks = nn.ModuleList([TimestepEmbedSequential(nn.Conv2d(in_channels, model_channels, 3, padding=1))])
        chans = [model_channels]
        ch, ds = model_channels, 1
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                layers: List[nn.Module] = [ResBlock(ch, ted, dropout, out_channels=mult * model_channels)]
                ch = mult * model_channels
                if ds in attention_resolutions:
                    layers.append(attn_layer(ch))
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                chans.append(ch)
            if level != len(channel_mult) - 1:
                self.input_blocks.append(TimestepEmbedSequential(Downsample(ch, conv_resample, out_channels=ch)))
                chans.append(ch)
                ds *= 2
        self.middle_block = TimestepEmbedSequential(ResBlock(ch, ted, dropout), attn_layer(ch), ResBlock(ch, ted, dropout))
        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                ich = chans.pop()
                layers = [ResBlock(ch + ich, ted, dropout, out_channels=model_channels * mult)]
                ch = model_channels * mult
                if ds in attention_resolutions:
                    layers.append(attn_layer(ch))
                if level and i == num_res_blocks:
                    layers.append(Upsample(ch, conv_resample, out_channels=ch))
                    ds //= 2
                self.output_blocks.append(TimestepEmbedSequential(*layers))
        self.out = nn.Sequential(nn.GroupNorm(32, ch), nn.SiLU(),
                                 zero_module(nn.Conv2d(model_channels, out_channels, 3, padding=1)))

        self._compute_dtype = {"bf16": torch.bfloat16, "fp32": torch.float32}[compute_dtype]
        self._engine: Optional["UNetEngine"] = None
        self.use_cuda_graph = False   # see enable_cuda_graph
        self.grad_ready_hook = None   # set by parallel.GradBuckets (data-parallel training)
        self._graphs = {}
        self._loop_scope = False      # see sampling_scope / cfg_pair
        self._pair_hint = False
        self._ctx_seen = None
        self.register_load_state_dict_post_hook(lambda m, keys: m.invalidate_packed())

    # ---- engine management -------------------------------------------------------------------------
    def set_compute_dtype(self, name: str) -> "UNetModel":
        self._compute_dtype = {"bf16": torch.bfloat16, "fp32": torch.float32}[name]
        self._engine = None
        return self

    def invalidate_packed(self):
        """Drop the packed kernel-layout weights (call after modifying parameters in place)."""
        self._engine = None
        self._graphs = {}
        self._ctx_seen = None

    # The packed weights (and the CUDA graphs over them) are a cache of the parameters.  `load_state_dict`, `_apply`,
    # optim.FusedAdamWEMA and ema.LitEma invalidate it explicitly; every eval forward additionally compares the
    # parameters' autograd version counters (bumped by torch optimizers and any in-place op; no device sync), and
    # `revalidate_packed()` -- called by the samplers once per `sample()` -- compares a content probe, which also
    # catches writes through `.data` (the reference's LitEma.copy_to, ema.py:46-53).
    def _version_fingerprint(self) -> int:
        return sum(p._version for p in self._plist)

    def _content_probe(self) -> torch.Tensor:
        with torch.no_grad():
            return torch.stack([p.detach().reshape(-1)[0].float() for p in self._plist[::8]])   # views + ONE kernel

    def revalidate_packed(self) -> bool:
        """True if the cached engine still matches the parameters; otherwise drops it (one device sync)."""
        if self._engine is None or torch.cuda.is_current_stream_capturing():
            return True
        if self._version_fingerprint() == self._engine_fp and torch.equal(self._content_probe(), self._engine_probe):
            return True
        self.invalidate_packed()
        return False

    def _get_engine(self) -> "UNetEngine":
        if self._engine is not None and self._version_fingerprint() != self._engine_fp:
            self.invalidate_packed()
        if self._engine is None:
            self._plist = list(self.parameters())
            self._engine = UNetEngine(self, self._compute_dtype)
            self._engine_fp = self._version_fingerprint()
            if not torch.cuda.is_current_stream_capturing():
                self._engine_probe = self._content_probe()
        return self._engine

    def enable_cuda_graph(self, flag: bool = True) -> "UNetModel":
        """Replay the forward pass (~300 kernel launches) as ONE CUDA graph per input shape.  The
        launch sequence has no host synchronisation or data-dependent control flow, so capture is
        exact; inputs are copied into static buffers and the result is returned as a fresh tensor."""
        self.use_cuda_graph = flag
        if not flag:
            self._graphs = {}
        return self

    def _apply(self, fn, *a, **k):
        self._engine = None
        self._graphs = {}
        return super()._apply(fn, *a, **k)

    # ---- hints from the samplers -----------------------------------------------------------------------
    # (attributes of the module, not thread-local: one sampling loop per model at a time)
    @contextlib.contextmanager
    def sampling_scope(self):
        """Entered by the samplers around one denoising loop: inside it a context tensor that is the SAME object (and
        autograd version) as in the previous call is not projected again -- the K / V rows and the collapsed
        cross-attention operands depend on the context and the weights only (ddim.py:118-131 passes one `cond` to
        every step).  Outside the scope every forward projects its context."""
        prev = self._loop_scope
        self._loop_scope = True
        try:
            yield self
        finally:
            self._loop_scope = prev
            if not prev:
                self._ctx_seen = None
                for ent in self._graphs.values():
                    ent["ctx_ref"] = None

    @contextlib.contextmanager
    def cfg_pair(self):
        """Entered by the samplers around `apply_model(torch.cat([x] * 2), torch.cat([t] * 2), cat([uc, c]))`
        (ddim.py:176-179): both halves of the batch carry the same images and timesteps, so the layers in front of
        the first cross-attention are computed once (UNetEngine.forward(shared_halves=True))."""
        prev = self._pair_hint
        self._pair_hint = True
        try:
            yield self
        finally:
            self._pair_hint = prev

    def _projection(self, engine, context, holder: dict, static_ctx=None):
        """The context projection for this call: reused inside a sampling_scope when `context` is the tensor seen last."""
        same = (self._loop_scope and holder.get("ctx_ref") is context and holder.get("ctx_ver") == context._version
                and not os.environ.get("EALDM_NO_CTX_REUSE"))
        if not same:
            src = context
            if static_ctx is not None:
                static_ctx.copy_(context)
                src = static_ctx
            holder["proj"] = engine.project_context(src, into=holder.get("proj") if static_ctx is not None else None)
            holder["ctx_ref"] = context if self._loop_scope else None
            holder["ctx_ver"] = context._version
        return holder["proj"]

    def _forward_graphed(self, x, timesteps, context, pair: bool):
        key = (tuple(x.shape), None if context is None else tuple(context.shape), self._compute_dtype, pair)
        ent = self._graphs.get(key)
        if ent is None:
            sx = x.detach().float().contiguous().clone()
            st = timesteps.detach().to(device=x.device, dtype=torch.int64).contiguous().clone()
            sc = None if context is None else context.detach().float().contiguous().clone()
            ent = {"sx": sx, "st": st, "sc": sc, "ctx_ref": None}
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):       # warm-up outside capture (lazy attributes, allocator)
                proj = None if sc is None else self._engine.project_context(sc)
                self._engine.forward(sx, st, sc, ctx_proj=proj, shared_halves=pair)
            torch.cuda.current_stream().wait_stream(side)
            ent["proj"] = proj      # static: refilled outside the graph whenever the context changes
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                ent["sy"] = self._engine.forward(sx, st, sc, ctx_proj=proj, shared_halves=pair)
            ent["graph"] = graph
            self._graphs[key] = ent
        ent["sx"].copy_(x)
        ent["st"].copy_(timesteps)
        if context is not None:
            self._projection(self._engine, context, ent, static_ctx=ent["sc"])
        ent["graph"].replay()
        return ent["sy"].clone()

    def convert_to_fp16(self):  # reference stubs (openaimodel.py:694-708): no-ops there as well
        pass

    def convert_to_fp32(self):
        pass

    def forward(self, x, timesteps=None, context=None, y=None, **kwargs):
        """x [N, C, H, W] fp32 (CUDA), timesteps [N] int64, context [N, T, context_dim] -> [N, out_channels, H, W]."""
        assert (y is not None) == (self.num_classes is not None), \
            "must specify y if and only if the model is class-conditional"
        if not x.is_cuda:
            raise RuntimeError("ealdm_b200.UNetModel runs on CUDA (sm_100a) only; there is no CPU fallback")
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            # training step: forward with saved activations + hand-written backward (train.py)
            from .train import unet_forward_train
            return unet_forward_train(self, x, timesteps, context)
        engine = self._get_engine()
        pair = self._pair_hint and x.shape[0] % 2 == 0
        if pair and os.environ.get("EALDM_CHECK_CFG_PAIR"):   # debugging aid (one device sync): the caller's guarantee
            hb = x.shape[0] // 2
            assert torch.equal(x[:hb], x[hb:]) and torch.equal(timesteps[:hb], timesteps[hb:]), \
                "cfg_pair(): the two halves of the batch must hold the same images and timesteps"
        if self.use_cuda_graph and not torch.cuda.is_current_stream_capturing():
            return self._forward_graphed(x, timesteps, context, pair)
        proj = None
        if context is not None and self._loop_scope and not torch.cuda.is_current_stream_capturing():
            if self._ctx_seen is None:
                self._ctx_seen = {}
            proj = self._projection(engine, context, self._ctx_seen)
        return engine.forward(x, timesteps, context, ctx_proj=proj, shared_halves=pair)


# ---- execution engine --------------------------------------------------------------------------------
class _PackedConv:
    __slots__ = ("w", "b", "cout")

    def __init__(self, w, b, cout):
        self.w, self.b, self.cout = w, b, cout


class UNetEngine:
    """Packs a UNetModel's parameters to kernel layout and runs the forward pass as a straight-line
    sequence of libealdm_b200 launches on the current CUDA stream (CUDA-graph capturable: no host
    synchronisation, no data-dependent control flow)."""
    _fused_geglu = True   # the training engine keeps the GEGLU pre-activation instead (train.py)
    # LayerNorm folded into the GEMMs around it (producer row statistics + consumer epilogue correction): built, parity
    # tested and measured NOT faster than the streaming LayerNorm passes it removes (same-box A/B 63.3 vs 64.0 samples/s:
    # the producers pay for a bf16 shadow + statistics, the GEGLU epilogue for one more FMA and a shared-memory operand
    # per element), hence opt-in
    _fold_ln = bool(os.environ.get("EALDM_LN_FOLD"))
    # one-kernel GEGLU FeedForward (csrc/ff_fused.cu): bit-identical to the two GEMMs and measured NOT faster yet
    # (269 us against 140 + 95 us at level 0), hence opt-in
    _fused_ff = bool(os.environ.get("EALDM_FUSED_FF"))
    _phased_upsample = not os.environ.get("EALDM_NO_PHASED_UPSAMPLE")   # A/B switch; the training engine saves `up`
    # cross-attention collapsed onto the context (packing.collapse_cross_attention): to_q -> 4-key attention -> to_out
    # become two per-image GEMMs of K = C / N = 32 and K = 32 / N = C; A/B switch, the training engine keeps q / k / v
    _collapse_xattn = not os.environ.get("EALDM_NO_XATTN_COLLAPSE")
    # ... also at the 8x8 level (64 tokens per image, two images per tile).  There the projection of the context by
    # heads * C = 32768 columns per layer costs more than the 64 tokens of an image save -- unless it is made once per
    # sampling loop (UNetModel.sampling_scope), which the samplers do: -0.3 ms per DDIM step for 0.25 ms per loop and
    # 0.8 GB of projection weights / operands; a stand-alone forward pays the 0.25 ms every time.  A/B switch
    _collapse_xattn_8x8 = not os.environ.get("EALDM_NO_XATTN_COLLAPSE_8X8")
    # LayerNorm applied by the epilogue of the GEMM that produces its input (width 256: a CTA holds whole rows); A/B switch
    _ln_epilogue = not os.environ.get("EALDM_NO_LN_EPILOGUE")
    # GroupNorm (+ SiLU) applied by the epilogue of the conv that produces its input (the ResBlock's second GroupNorm:
    # conv1's result is never written un-normalised; the GroupNorm in front of a transformer block: second output of
    # the ResBlock's conv2); A/B switch
    _gn_epilogue = not os.environ.get("EALDM_NO_GN_EPILOGUE")

    def __init__(self, m: UNetModel, dtype: torch.dtype):
        L.load()
        self.m, self.dt = m, dtype
        dev = next(m.parameters()).device
        assert dev.type == "cuda", "move the model to CUDA before the first forward"
        self.dev = dev
        f32 = lambda t: t.detach().float().contiguous()  # noqa: E731
        self._cast_cache = {}     # parameter -> compute-dtype copy (shared by the forward and dgrad packings)
        pk = lambda conv: pack_conv_weight(self._c(conv.weight), dtype)  # noqa: E731
        half = m.model_channels // 2
        # computed on the host exactly as the reference does (util.py:160-162), cached on the module so that
        # building an engine issues no host->device copy (engines are rebuilt inside CUDA-graph captures)
        fr = getattr(m, "_freqs_cache", None)
        if fr is None or fr.device != dev:
            fr = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32) / half).to(dev)
            m._freqs_cache = fr
        self.freqs = fr
        self.te0 = _PackedConv(self._c(m.time_embed[0].weight), f32(m.time_embed[0].bias), 0)
        self.te2 = _PackedConv(self._c(m.time_embed[2].weight), f32(m.time_embed[2].bias), 0)

        self.blocks = []          # execution list
        emb_w, emb_b = [], []     # all ResBlock emb_layers stacked into one GEMM
        kv_w = []                 # all cross-attention to_k / to_v stacked into one GEMM
        xc_w = []                 # collapsed cross-attention: all [G; H] context projections stacked into one GEMM
        self.emb_cols = 0
        self.kv_cols = 0
        self.xc_cols = 0

        def pack_res(rb: ResBlock):
            d = {"kind": "res", "cin": rb.channels, "cout": rb.out_channels}
            d["gn1"] = (f32(rb.in_layers[0].weight), f32(rb.in_layers[0].bias))
            d["conv1"] = _PackedConv(pk(rb.in_layers[2]), f32(rb.in_layers[2].bias), rb.out_channels)
            d["emb_col0"] = self.emb_cols
            emb_w.append(self._c(rb.emb_layers[1].weight))
            emb_b.append(rb.emb_layers[1].bias.detach())
            self.emb_cols += rb.out_channels
            d["gn2"] = (f32(rb.out_layers[0].weight), f32(rb.out_layers[0].bias))
            w2, b2 = pk(rb.out_layers[3]), f32(rb.out_layers[3].bias)
            if isinstance(rb.skip_connection, nn.Identity):
                d["skip"] = False
            else:  # 1x1 skip conv accumulated in the same GEMM: K = 9*cout + cin
                w2 = torch.cat([w2, pk(rb.skip_connection)], dim=1).contiguous()
                b2 = b2 + f32(rb.skip_connection.bias)
                d["skip"] = True
            d["conv2"] = _PackedConv(w2, b2, rb.out_channels)
            return d

        def pack_st(st: SpatialTransformer):
            C_ = st.in_channels
            d = {"kind": "st", "c": C_, "heads": st.n_heads, "dh": st.d_head}
            assert st.d_head in (32, 64), "head_dim must be 32 or 64"
            d["norm"] = (f32(st.norm.weight), f32(st.norm.bias))
            d["proj_in"] = _PackedConv(pk(st.proj_in), f32(st.proj_in.bias), C_)
            d["proj_out"] = _PackedConv(pk(st.proj_out), f32(st.proj_out.bias), C_)
            d["blocks"] = []
            for tb in st.transformer_blocks:
                t = {}
                for i, ln in ((1, tb.norm1), (2, tb.norm2), (3, tb.norm3)):
                    t[f"ln{i}"] = (f32(ln.weight), f32(ln.bias))
                t["qkv"] = torch.cat([self._c(tb.attn1.to_q.weight), self._c(tb.attn1.to_k.weight),
                                      self._c(tb.attn1.to_v.weight)], dim=0)
                t["o1"] = _PackedConv(self._c(tb.attn1.to_out[0].weight),
                                      f32(tb.attn1.to_out[0].bias), C_)
                t["q2"] = self._c(tb.attn2.to_q.weight)
                t["kv_col0"] = self.kv_cols
                kv_w.append(torch.cat([self._c(tb.attn2.to_k.weight), self._c(tb.attn2.to_v.weight)], dim=0))
                self.kv_cols += 2 * C_
                t["o2"] = _PackedConv(self._c(tb.attn2.to_out[0].weight),
                                      f32(tb.attn2.to_out[0].bias), C_)
                # (only where a 128-row tile lies inside one image at the model's nominal resolution: the projection of
                # the context grows with heads * C = C^2 / 32 per layer, so unused blocks are not worth carrying)
                tok_ = (m.image_size // self._ds) ** 2
                if (self._collapse_xattn and self.dt == torch.bfloat16 and st.d_head == 32 and C_ % 64 == 0
                        and (tok_ % 128 == 0 or (tok_ == 64 and st.n_heads * 4 == 128 and self._collapse_xattn_8x8))
                        and st.n_heads * 4 <= 128):
                    g_, h_ = collapse_cross_attention(tb.attn2.to_q.weight, tb.attn2.to_k.weight, tb.attn2.to_v.weight,
                                                      tb.attn2.to_out[0].weight, st.n_heads, tb.attn2.scale, dtype)
                    t["xc_col0"] = self.xc_cols          # U block at xc_col0, Zt block at xc_col0 + heads * C
                    xc_w.extend([g_, h_])
                    self.xc_cols += 2 * st.n_heads * C_
                if self._fused_geglu:   # inference: value/gate rows interleaved for the fused GEGLU epilogue
                    wi, bi = geglu_interleave(self._c(tb.ff.net[0].proj.weight), tb.ff.net[0].proj.bias)
                    t["ff1"] = _PackedConv(wi, bi, 4 * C_)
                if self._fold_ln and self.dt == torch.bfloat16:
                    # LayerNorm folded into the GEMM behind it: LN(x) W^T + b = rstd (x (W.gamma)^T) - rstd mu c1 + c2
                    # with c1 = (W.gamma) 1 (from the bf16-rounded matrix the GEMM multiplies) and c2 = W beta + b
                    def fold(w, ln, bias=None):
                        w32 = w.detach().float()
                        wg = (w32 * ln.weight.detach().float()[None]).to(torch.bfloat16).contiguous()
                        c2 = w32 @ ln.bias.detach().float()
                        if bias is not None:
                            c2 = c2 + bias.detach().float()
                        return wg, wg.float().sum(1).contiguous(), c2.contiguous()
                    wqkv = torch.cat([tb.attn1.to_q.weight, tb.attn1.to_k.weight, tb.attn1.to_v.weight], dim=0)
                    t["qkv_ln"] = fold(wqkv, tb.norm1)
                    t["q2_ln"] = fold(tb.attn2.to_q.weight, tb.norm2)
                    wg, c1, c2 = fold(tb.ff.net[0].proj.weight, tb.norm3, tb.ff.net[0].proj.bias)
                    wgi, c2i = geglu_interleave(wg, c2)
                    _, c1i = geglu_interleave(wg, c1)
                    t["ff1_ln"] = (wgi, c1i, c2i)
                t["ff2"] = _PackedConv(self._c(tb.ff.net[2].weight), f32(tb.ff.net[2].bias), C_)
                d["blocks"].append(t)
            return d

        def pack_ab(ab: AttentionBlock):
            C_ = ab.channels
            assert C_ // ab.num_heads in (32, 64), "head_dim must be 32 or 64"
            return {"kind": "ab", "c": C_, "heads": ab.num_heads, "norm": (f32(ab.norm.weight), f32(ab.norm.bias)),
                    "qkv": _PackedConv(pk(ab.qkv), f32(ab.qkv.bias), 3 * C_),
                    "proj": _PackedConv(pk(ab.proj_out), f32(ab.proj_out.bias), C_)}

        def pack_layers(seq):
            out = []
            for layer in seq:
                if isinstance(layer, ResBlock):
                    out.append(pack_res(layer))
                elif isinstance(layer, SpatialTransformer):
                    out.append(pack_st(layer))
                elif isinstance(layer, AttentionBlock):
                    out.append(pack_ab(layer))
                elif isinstance(layer, Downsample):
                    out.append({"kind": "down", "c": layer.channels,
                                "conv": _PackedConv(pk(layer.op), f32(layer.op.bias), layer.out_channels)})
                    self._ds *= 2
                elif isinstance(layer, Upsample):
                    d = {"kind": "up", "c": layer.channels,
                         "conv": _PackedConv(pk(layer.conv), f32(layer.conv.bias), layer.out_channels)}
                    if self.dt == torch.bfloat16 and self._phased_upsample and layer.channels % 64 == 0:
                        # nearest-2x folded into the conv: four 2x2 output phases over the low-resolution input
                        d["w_phases"] = pack_upsample_phases(layer.conv.weight, dtype)
                    out.append(d)
                    self._ds //= 2
                elif isinstance(layer, nn.Conv2d):
                    d = {"kind": "conv_in", "conv": _PackedConv(pk(layer), f32(layer.bias), layer.out_channels)}
                    if self.dt == torch.bfloat16 and layer.in_channels < 64 and layer.kernel_size == (3, 3):
                        # tcgen05 needs 64-channel K blocks: the same taps over an input zero-padded to 64 channels
                        # (16x the useful FLOPs of a 4-channel conv and still 5x faster than the FFMA kernel)
                        w9 = d["conv"].w.reshape(layer.out_channels, 9, layer.in_channels)
                        wp = torch.zeros(layer.out_channels, 9, 64, dtype=w9.dtype, device=w9.device)
                        wp[:, :, :layer.in_channels] = w9
                        d["w_pad"] = wp.reshape(layer.out_channels, 9 * 64).contiguous()
                    out.append(d)
                else:
                    raise TypeError(type(layer))
            return out

        self._ds = 1              # downsampling factor of the level being packed (execution order)
        self.inp = [pack_layers(b) for b in m.input_blocks]
        self.mid = pack_layers(m.middle_block)
        self.outb = [pack_layers(b) for b in m.output_blocks]
        self.out_norm = (f32(m.out[0].weight), f32(m.out[0].bias))
        self.out_conv = _PackedConv(pk(m.out[2]), f32(m.out[2].bias), m.out_channels)
        self.emb_w = torch.cat(emb_w, dim=0)
        self.emb_b = torch.cat(emb_b, dim=0).float().contiguous()
        self.kv_w = torch.cat(kv_w, dim=0) if kv_w else None
        self.xc_w = torch.cat(xc_w, dim=0) if xc_w else None

        # channel bookkeeping for the zero-copy skip concatenation
        def block_out_channels(layers, cin):
            c = cin
            for d in layers:
                if d["kind"] in ("res",):
                    c = d["cout"]
                elif d["kind"] == "conv_in":
                    c = d["conv"].cout
            return c

        self._overlap_emb = (type(self) is UNetEngine and not os.environ.get("EALDM_NO_EMB_OVERLAP")
                             and len(self.inp) > 1 and self.inp[0][0]["kind"] == "conv_in")
        self._side = torch.cuda.Stream(device=dev) if self._overlap_emb else None
        self.skip_ch = []
        c = m.in_channels
        for layers in self.inp:
            c = block_out_channels(layers, c)
            self.skip_ch.append(c)
        self.mid_ch = c

    def _c(self, w: torch.Tensor) -> torch.Tensor:
        """Compute-dtype copy of a parameter, made once per engine."""
        t = self._cast_cache.get(id(w))
        if t is None:
            view = getattr(w, "_bf16_view", None)     # optim.FusedAdamWEMA keeps a bf16 copy of every parameter
            t = view if (view is not None and self.dt == torch.bfloat16) else w.detach().to(self.dt)
            self._cast_cache[id(w)] = t
        return t

    # ---- helpers --------------------------------------------------------------------------------------
    # The residual stream ("trunk": ResBlock / attention-block / resample outputs and the token stream
    # inside a transformer block) is kept in fp32, as it is under torch.autocast; only GEMM operands
    # are bf16.  A trunk tensor that is also a GEMM operand (1x1 skip conv, down/up-sampling conv) gets
    # a bf16 shadow written by the same epilogue (`out2`).  In fp32 mode master and shadow coincide.
    def _new(self, n, h, w, c, dtype=None) -> Act:
        return Act.empty(n, h, w, c, dtype or self.dt, self.dev)

    def _new_dual(self, n, h, w, c, gn: bool = True) -> "Dual":
        """A residual-stream tensor.  In bf16 mode its fp32 master carries a GroupNorm partial-statistics buffer
        that the producing tcgen05 epilogue fills, so the GroupNorm reading it is a single streaming pass."""
        f = self._new(n, h, w, c, torch.float32)
        if gn and self.dt == torch.bfloat16:
            f.with_gn_partial()
        return Dual(f, f if self.dt == torch.float32 else self._new(n, h, w, c))

    def _conv_first(self, d, x: "Dual", out: "Dual") -> "Dual":
        """input_blocks.0.  Inference in bf16 mode hands in an input whose buffer is zero-padded to 64 channels: the conv
        then runs on the tcgen05 kernel (statistics epilogue included).  Otherwise (fp32 parity mode, training engine)
        the 4-channel conv runs on the SIMT kernel, which has no statistics epilogue."""
        if "w_pad" in d and x.h.ld == 64 and x.h.c0 == 0 and x.h.dtype == torch.bfloat16:
            x64 = Act(x.h.buf, x.h.n, x.h.h, x.h.w, 64)
            ops.conv([ConvIn(x64, 3, 1, 1)], d["w_pad"], out.f, bias=d["conv"].b, out2=self._out2(out))
            return out
        plain = Act(out.f.buf, out.f.n, out.f.h, out.f.w, out.f.c, out.f.c0)
        ops.conv([ConvIn(x.h, 3, 1, 1)], d["conv"].w, plain, bias=d["conv"].b, out2=self._out2(out))
        if out.f.gp is not None:
            ops.gn_partial(out.f)
        return out

    def _out2(self, d: "Dual"):
        return None if d.h is d.f else d.h

    # largest image (pixels) for the two variants: conv1 -> GroupNorm -> SiLU (the un-normalised tensor is never written)
    # pays at every level; conv2 + the transformer block's GroupNorm as a second output pays where an image lies in ONE
    # tile (16x16, 8x8) and loses 7 us per launch at 32x32, where four CTA pairs wait for each other
    _gna_maxhw = (int(os.environ.get("EALDM_GNA_A_MAXHW", "1024")), int(os.environ.get("EALDM_GNA_B_MAXHW", "256")))

    def _gn_epilogue_ok(self, c: int, h: int, w: int, variant: int = 0) -> bool:
        """GroupNorm32 over c channels can run in the producing conv's epilogue (ops.conv(gn_apply=...))."""
        return (self._gn_epilogue and self.dt == torch.bfloat16 and type(self) is UNetEngine and c >= 256
                and c % 32 == 0 and c // 32 in (8, 16, 32) and (h * w) % 32 == 0 and h * w <= self._gna_maxhw[variant])

    def _res(self, d, x: "Dual", emb_all: torch.Tensor, dest: Optional["Dual"], shadow: bool = True,
             norm_next: Optional[tuple] = None) -> "Dual":
        """norm_next = (gamma, beta, eps) of the GroupNorm (without SiLU) the NEXT layer starts with: conv2's epilogue
        then also writes it (result.f.gny)."""
        n, h, w = x.f.n, x.f.h, x.f.w
        hn = self._new(n, h, w, x.f.c)
        ops.group_norm(x.f, d["gn1"][0], d["gn1"][1], 1e-5, hn, self.stats, silu=True)
        hn2 = self._new(n, h, w, d["cout"])
        gna = self._gn_epilogue_ok(d["cout"], h, w) and hn2.with_gn_partial().gp is not None
        if gna:
            # conv1 + emb -> GroupNorm -> SiLU in one kernel: the fp32 intermediate h1 is never written
            ops.conv([ConvIn(hn, 3, 1, 1)], d["conv1"].w, hn2, bias=d["conv1"].b, rowvec=emb_all,
                     rowvec_col0=d["emb_col0"], gn_apply=(d["gn2"][0], d["gn2"][1], 1e-5, 32, True, True))
        else:
            hn2.gp = None
            h1 = self._new(n, h, w, d["cout"], torch.float32)   # bf16 here costs ~40 % of the 1e-2 eps budget (measured)
            if self.dt == torch.bfloat16:
                h1.with_gn_partial()
            ops.conv([ConvIn(hn, 3, 1, 1)], d["conv1"].w, h1, bias=d["conv1"].b, rowvec=emb_all,
                     rowvec_col0=d["emb_col0"])
            ops.group_norm(h1, d["gn2"][0], d["gn2"][1], 1e-5, hn2, self.stats, silu=True)
        if dest is not None:
            out = dest
        elif shadow:
            out = self._new_dual(n, h, w, d["cout"])
        else:   # only the fp32 master is read downstream (GroupNorm + residual of a SpatialTransformer)
            f = self._new(n, h, w, d["cout"], torch.float32)
            if self.dt == torch.bfloat16:
                f.with_gn_partial()
            out = Dual(f, f)
        kw = {"out2": self._out2(out)}
        if (norm_next is not None and kw["out2"] is None and out.f.gp is not None and out.f.gunit == 8
                and self._gn_epilogue_ok(d["cout"], h, w, 1)):
            out.f.gny = self._new(n, h, w, d["cout"])
            kw = {"out2": out.f.gny, "gn_apply": (norm_next[0], norm_next[1], norm_next[2], 32, False, False)}
        if d["skip"]:
            ops.conv([ConvIn(hn2, 3, 1, 1), ConvIn(x.h, 1, 1, 0)], d["conv2"].w, out.f, bias=d["conv2"].b, **kw)
        else:
            ops.conv([ConvIn(hn2, 3, 1, 1)], d["conv2"].w, out.f, bias=d["conv2"].b, residual=x.f, **kw)
        return out

    def _st(self, d, x: "Dual", kv_all: Optional[Act], n_ctx: int, dest: Optional["Dual"],
            share_full: Optional["Dual"] = None) -> "Dual":
        """share_full (classifier-free-guidance pair, see forward): `x` is the first-half view of `share_full`; up to
        the first cross-attention the block runs on that half, then the token stream is duplicated."""
        C_, heads, dh = d["c"], d["heads"], d["dh"]
        n, h, w = x.f.n, x.f.h, x.f.w
        tok = h * w
        f32 = torch.float32
        xn = x.f.gny          # written by the producing conv's epilogue (see _res)
        if xn is None:
            xn = self._new(n, h, w, C_)
            ops.group_norm(x.f, d["norm"][0], d["norm"][1], 1e-6, xn, self.stats, silu=False)
        if kv_all is None:
            raise RuntimeError("SpatialTransformer needs a context tensor")
        fold = "qkv_ln" in d["blocks"][0]     # bf16 inference engine: no stand-alone LayerNorm passes (see pack_st)
        fused_ff = self._fused_ff and C_ == 256 and self.dt == torch.bfloat16
        nb = len(d["blocks"])

        def normed(src: Act, src_h: Optional[Act], tb, i: int, key: str, w_plain, n_out: int, **kw) -> Act:
            """LayerNorm i of the block followed by the linear layer `key`: folded into that GEMM (which then reads
            the bf16 shadow of the raw stream and the producer's row statistics) or as a LayerNorm pass + GEMM."""
            y = self._new(n, h, w, n_out // 2 if kw.get("act") == L.ACT_GEGLU else n_out)
            if src.lny is not None:       # the producing GEMM's epilogue already wrote LayerNorm(src)
                ops.linear(src.lny, w_plain[0], y, bias=w_plain[1], **kw)
            elif fold and src.ln is not None:
                wg, c1, c2 = tb[key + "_ln"]
                ops.linear(src_h, wg, y, bias=c2, ln=(src.ln, c1, C_, 1e-5), **kw)
            else:
                a = self._new(n, h, w, C_)
                ops.layer_norm(src, tb[f"ln{i}"][0], tb[f"ln{i}"][1], 1e-5, a)
                ops.linear(a, w_plain[0], y, bias=w_plain[1], **kw)
            return y

        def stream(last: bool = False):
            """A token-stream tensor: fp32 master (+ its bf16 shadow when a folded LayerNorm GEMM will read it)."""
            f = self._new(n, h, w, C_, None if last else f32)
            return f, (self._new(n, h, w, C_) if (fold and not last) else None)

        ln_epi = self._ln_epilogue and not fold and C_ == 256 and self.dt == torch.bfloat16

        def ln_out(dst: Act, tb, i: int) -> dict:
            """Keyword arguments that make the GEMM producing `dst` also write LayerNorm i of block `tb` of it."""
            if not ln_epi or tb is None:
                return {}
            dst.lny = self._new(n, h, w, C_)
            return {"out2": dst.lny, "ln_apply": (tb[f"ln{i}"][0], tb[f"ln{i}"][1], 1e-5)}

        t, t_h = stream()
        if ln_epi:
            ops.linear(xn, d["proj_in"].w, t, bias=d["proj_in"].b, **ln_out(t, d["blocks"][0], 1))
        else:
            ops.linear(xn, d["proj_in"].w, t, bias=d["proj_in"].b, out2=t_h, ln_stats=fold)
        for bi, tb in enumerate(d["blocks"]):
            qkv = normed(t, t_h, tb, 1, "qkv", (tb["qkv"], None), 3 * C_)
            o = self._new(n, h, w, C_)
            ops.attention(qkv.cols(0, C_), qkv.cols(C_, C_), qkv.cols(2 * C_, C_), o, batch=n, heads=heads,
                          head_dim=dh, n_q=tok, n_kv=tok, scale=dh ** -0.5)
            t1, t1_h = stream()
            if share_full is not None and bi == 0:
                # everything so far saw x and t only: the second half of the batch is a copy of the first
                nf = share_full.f.n
                t1 = self._new(nf, h, w, C_, f32)
                t1v = t1.images(0, n)
                kw = {}
                if ln_epi:
                    t1.lny = self._new(nf, h, w, C_)
                    t1v.lny = t1.lny.images(0, n)
                    kw = {"out2": t1v.lny, "ln_apply": (tb["ln2"][0], tb["ln2"][1], 1e-5)}
                ops.linear(o, tb["o1"].w, t1v, bias=tb["o1"].b, residual=t, **kw)
                for a in (t1, t1.lny, share_full.f):
                    if a is not None:
                        _dup_first_half(a)
                x, n = share_full, nf
            elif ln_epi:
                ops.linear(o, tb["o1"].w, t1, bias=tb["o1"].b, residual=t, **ln_out(t1, tb, 2))
            else:
                ops.linear(o, tb["o1"].w, t1, bias=tb["o1"].b, residual=t, out2=t1_h, ln_stats=fold)
            ff_fold = fold and not fused_ff
            t2 = self._new(n, h, w, C_, f32)
            t2_h = self._new(n, h, w, C_) if ff_fold else None
            xc = getattr(self, "xc_all", None)
            two = tok == 64 and heads * n_ctx == 128 and h == 8 and w == 8     # two images per 128-row tile
            if "xc_col0" in tb and xc is not None and not fold and (tok % 128 == 0 or two) and heads * n_ctx <= 128:
                # collapsed cross-attention: logits = LN2(t1) U_n^T (softmax over the 4 keys in the epilogue), then
                # t2 = P Zt_n + bias + t1; U_n / Zt_n are column windows of the context projection xc_all
                a2 = t1.lny
                if a2 is None:
                    a2 = self._new(n, h, w, C_)
                    ops.layer_norm(t1, tb["ln2"][0], tb["ln2"][1], 1e-5, a2)
                pr = self._new(n, h, w, heads * n_ctx * (2 if two else 1))
                ops.conv([ConvIn(a2)], xc.buf, pr, act=L.ACT_SOFTMAX4, wimg=(tb["xc_col0"], n_ctx, heads, C_))
                ops.conv([ConvIn(pr)], xc.buf, t2, bias=tb["o2"].b, residual=t1, adjoint=True,
                         wimg=(tb["xc_col0"] + heads * C_, n_ctx, heads, C_), **ln_out(t2, tb, 3))
            else:
                q2 = normed(t1, t1_h, tb, 2, "q2", (tb["q2"], None), C_)
                o2 = self._new(n, h, w, C_)
                kc = tb["kv_col0"]
                ops.attention(q2, kv_all.cols(kc, C_), kv_all.cols(kc + C_, C_), o2, batch=n, heads=heads, head_dim=dh,
                              n_q=tok, n_kv=n_ctx, scale=dh ** -0.5)
                if ln_epi:
                    ops.linear(o2, tb["o2"].w, t2, bias=tb["o2"].b, residual=t1, **ln_out(t2, tb, 3))
                else:
                    ops.linear(o2, tb["o2"].w, t2, bias=tb["o2"].b, residual=t1, out2=t2_h, ln_stats=ff_fold)
            last = bi == nb - 1   # the last t feeds proj_out as a GEMM operand -> compute dtype
            t, t_h = stream(last)
            if fused_ff:
                # FF1 -> GEGLU -> FF2 -> + residual in one kernel: the 8C / 4C wide intermediates stay on the SM
                a = self._new(n, h, w, C_)
                ops.layer_norm(t2, tb["ln3"][0], tb["ln3"][1], 1e-5, a)
                ops.ff_geglu_fused(a, tb["ff1"].w, tb["ff1"].b, tb["ff2"].w, tb["ff2"].b, t2, t)
            else:
                gg = normed(t2, t2_h, tb, 3, "ff1", (tb["ff1"].w, tb["ff1"].b), 8 * C_, act=L.ACT_GEGLU)
                if ln_epi and not last:
                    ops.linear(gg, tb["ff2"].w, t, bias=tb["ff2"].b, residual=t2, **ln_out(t, d["blocks"][bi + 1], 1))
                else:
                    ops.linear(gg, tb["ff2"].w, t, bias=tb["ff2"].b, residual=t2, out2=t_h, ln_stats=fold and not last)
        out = dest if dest is not None else self._new_dual(n, h, w, C_)
        ops.linear(t, d["proj_out"].w, out.f, bias=d["proj_out"].b, residual=x.f, out2=self._out2(out))
        return out

    def _ab(self, d, x: "Dual", dest: Optional["Dual"]) -> "Dual":
        C_, heads = d["c"], d["heads"]
        dh = C_ // heads
        n, h, w = x.f.n, x.f.h, x.f.w
        xn = x.f.gny
        if xn is None:
            xn = self._new(n, h, w, C_)
            ops.group_norm(x.f, d["norm"][0], d["norm"][1], 1e-5, xn, self.stats, silu=False)
        qkv = self._new(n, h, w, 3 * C_)
        ops.linear(xn, d["qkv"].w, qkv, bias=d["qkv"].b)
        o = self._new(n, h, w, C_)
        span = 3 * C_ - 2 * dh  # column window that keeps every head's slice inside the buffer
        ops.attention(qkv.cols(0, span), qkv.cols(dh, span), qkv.cols(2 * dh, span), o, batch=n, heads=heads,
                      head_dim=dh, n_q=h * w, n_kv=h * w, scale=dh ** -0.5, head_stride_q=3 * dh,
                      head_stride_kv=3 * dh)
        out = dest if dest is not None else self._new_dual(n, h, w, C_)
        ops.linear(o, d["proj"].w, out.f, bias=d["proj"].b, residual=x.f, out2=self._out2(out))
        return out

    def _run(self, layers, x: "Dual", emb_all, kv_all, n_ctx, dest: Optional["Dual"]) -> "Dual":
        for i, d in enumerate(layers):
            dst = dest if i == len(layers) - 1 else None
            k = d["kind"]
            n, h, w = x.f.n, x.f.h, x.f.w
            if k == "res":
                # the bf16 operand shadow of the result is written only if the next layer reads it (a transformer
                # block reads the fp32 master alone)
                nxt = layers[i + 1] if i + 1 < len(layers) else None
                nk = None if nxt is None else nxt["kind"]
                nn_ = None if nk not in ("st", "ab") else (nxt["norm"][0], nxt["norm"][1], 1e-6 if nk == "st" else 1e-5)
                x = self._res(d, x, emb_all, dst, shadow=nk not in ("st", "ab"), norm_next=nn_)
            elif k == "st":
                x = self._st(d, x, kv_all, n_ctx, dst)
            elif k == "ab":
                x = self._ab(d, x, dst)
            elif k == "conv_in":
                out = dst if dst is not None else self._new_dual(n, h, w, d["conv"].cout)
                x = self._conv_first(d, x, out)
            elif k == "down":
                out = dst if dst is not None else self._new_dual(n, h // 2, w // 2, d["conv"].cout)
                ops.conv([ConvIn(x.h, 3, 2, 1)], d["conv"].w, out.f, bias=d["conv"].b, out2=self._out2(out))
                x = out
            elif k == "up":
                out = dst if dst is not None else self._new_dual(n, h * 2, w * 2, d["conv"].cout)
                if "w_phases" in d and h * w >= 32 and (h & (h - 1)) == 0 and (w & (w - 1)) == 0:
                    ops.conv([ConvIn(x.h, 3, 1, 1, upsample=1)], d["w_phases"], out.f, bias=d["conv"].b,
                             out2=self._out2(out), upsample_phases=True)
                elif self.dt == torch.bfloat16:
                    up = self._new(n, h * 2, w * 2, x.h.c)
                    ops.upsample_nearest2x(x.h, up)
                    ops.conv([ConvIn(up, 3, 1, 1)], d["conv"].w, out.f, bias=d["conv"].b, out2=self._out2(out))
                else:
                    ops.conv([ConvIn(x.h, 3, 1, 1, upsample=1)], d["conv"].w, out.f, bias=d["conv"].b)
                x = out
            else:
                raise ValueError(k)
        return x

    def _shareable(self, n: int) -> bool:
        """The classifier-free-guidance prefix is shared when the first block after the input conv is
        [ResBlock, SpatialTransformer] (every shipped config) and LayerNorm is not folded into the GEMMs."""
        k = [[d["kind"] for d in layers] for layers in self.inp[:2]]
        return (n % 2 == 0 and n >= 2 and self.dt == torch.bfloat16 and type(self) is UNetEngine and not self._fold_ln
                and not os.environ.get("EALDM_NO_CFG_SHARE") and k == [["conv_in"], ["res", "st"]])

    def _run_shared(self, layers, x: "Dual", emb_all, kv_all, n_ctx, dest: "Dual") -> "Dual":
        """input_blocks.1 of a guidance pair: ResBlock, GroupNorm, proj_in and the first self-attention run on the
        first half of the batch (the second half holds the same images and timesteps; only the context differs)."""
        n2 = x.f.n // 2
        xh = Dual(x.f.images(0, n2), x.h.images(0, n2))
        rf = self._new(x.f.n, x.f.h, x.f.w, layers[0]["cout"], torch.float32).with_gn_partial()
        r = Dual(rf, rf)      # (no bf16 shadow: the transformer block reads the fp32 master alone)
        rv = rf.images(0, n2)
        rh = Dual(rv, rv)
        self._res(layers[0], xh, emb_all[:n2], rh, shadow=False,
                  norm_next=(layers[1]["norm"][0], layers[1]["norm"][1], 1e-6))
        return self._st(layers[1], rh, kv_all, n_ctx, dest, share_full=r)

    # ---- forward ----------------------------------------------------------------------------------------
    @torch.no_grad()
    def project_context(self, context: torch.Tensor, into: Optional[tuple] = None) -> tuple:
        """(kv_all, xc_all): the K / V rows of every cross-attention layer and the collapsed-attention operands
        [G; H] x context of the layers that use them, each as ONE GEMM over the context.  `into`: buffers of an
        earlier call to refill (the CUDA graph of the forward reads them)."""
        dt, dev = self.dt, self.dev
        n, n_ctx = context.shape[0], context.shape[1]
        csrc = Act(context.float().reshape(n * n_ctx, -1).contiguous(), n, 1, n_ctx)
        ctx = ops.copy2d(csrc, Act.empty(n, 1, n_ctx, csrc.c, dt, dev)) if dt != torch.float32 else csrc
        kv_all = into[0] if into is not None else Act.empty(n, 1, n_ctx, self.kv_cols, dt, dev)
        ops.linear(ctx, self.kv_w, kv_all)
        xc_all = None
        if self.xc_w is not None and n_ctx == 4:
            xc_all = into[1] if into is not None else Act.empty(n, 1, n_ctx, self.xc_cols, dt, dev)
            ops.linear(ctx, self.xc_w, xc_all)
        return kv_all, xc_all

    def forward(self, x: torch.Tensor, timesteps: torch.Tensor, context: Optional[torch.Tensor],
                ctx_proj: Optional[tuple] = None, shared_halves: bool = False) -> torch.Tensor:
        """ctx_proj: `project_context(context)` made earlier (the context is loop-invariant in a sampling run).
        shared_halves: the caller guarantees x[:n/2] == x[n/2:] and timesteps[:n/2] == timesteps[n/2:] (classifier-free
        guidance: `torch.cat([x] * 2)`, ddim.py:176-178) -- everything in front of the first cross-attention then
        depends on the first half only and is computed once (bit-identical results, see _run_shared)."""
        m, dt, dev = self.m, self.dt, self.dev
        n, cin, H, W = x.shape
        assert cin == m.in_channels
        x = x.float().contiguous()
        timesteps = timesteps.to(device=dev, dtype=torch.int64).contiguous()
        assert timesteps.shape == (n,)
        # one GroupNorm scratch for the whole forward, sized for the widest concat at the finest level
        self.stats = ops.group_norm_workspace(n, H * W, 2 * self.mid_ch, dev)

        # timestep embedding -> time_embed MLP -> SiLU(emb) -> every ResBlock's emb_layers in ONE GEMM.  Four tiny
        # launches (M = n rows) that nothing needs before the first ResBlock: they run on a side stream next to the
        # input convolution (a fork / join the CUDA graph keeps).  Buffers are allocated on the main stream and stay
        # referenced until the join, so the caching allocator cannot hand them out early.
        temb = torch.empty((n, m.model_channels), dtype=dt, device=dev)
        ted = self.te0.w.shape[0]
        e1 = Act.empty(1, 1, n, ted, dt, dev)
        semb = Act.empty(1, 1, n, ted, dt, dev)
        emb_all = Act.empty(1, 1, n, self.emb_cols, torch.float32, dev)
        main = torch.cuda.current_stream()
        side = self._side if self._overlap_emb else main
        if side is not main:
            side.wait_stream(main)
        with torch.cuda.stream(side):
            ops.timestep_embedding(timesteps, m.model_channels, self.freqs, temb)
            ops.linear(Act(temb, 1, 1, n), self.te0.w, e1, bias=self.te0.b, act=L.ACT_SILU)
            ops.linear(e1, self.te2.w, semb, bias=self.te2.b, act=L.ACT_SILU)
            ops.linear(semb, self.emb_w, emb_all, bias=self.emb_b)
        emb_ready = side.record_event() if side is not main else None

        # context -> K/V of every cross-attention layer in ONE GEMM (loop-invariant across DDIM steps: the samplers
        # hand in the projection they made once per `sample()`, see UNetModel.sampling_scope)
        kv_all, n_ctx = None, 0
        if self.kv_w is not None:
            if context is None:
                raise RuntimeError("this UNet was built with context_dim: pass context=[N, T, context_dim]")
            assert context.shape[0] == n and context.shape[2] == self.kv_w.shape[1]
            n_ctx = context.shape[1]
            kv_all, self.xc_all = ctx_proj if ctx_proj is not None else self.project_context(context)

        # concat buffers of the output blocks: [h | skip]; producers write straight into them
        n_in = len(self.inp)
        res = [(H, W)]
        for layers in self.inp[1:]:
            hh, ww = res[-1]
            res.append((hh // 2, ww // 2) if layers[0]["kind"] == "down" else (hh, ww))
        cat: List[Dual] = []
        h_ch = self.mid_ch
        for j, layers in enumerate(self.outb):
            i = n_in - 1 - j                       # skip partner (hs.pop(), openaimodel.py:736)
            hh, ww = res[i]
            cat.append(self._new_dual(n, hh, ww, h_ch + self.skip_ch[i]))
            h_ch = layers[0]["cout"]

        def window(j, c0, c):
            return Dual(cat[j].f.cols(c0, c), cat[j].f.cols(c0, c) if cat[j].h is cat[j].f else cat[j].h.cols(c0, c))

        skip_dst = [window(n_in - 1 - i, cat[n_in - 1 - i].f.c - self.skip_ch[i], self.skip_ch[i]) for i in range(n_in)]

        if self.dt == torch.bfloat16 and cin < 64 and type(self) is UNetEngine:
            xin = Act(torch.zeros((n * H * W, 64), dtype=self.dt, device=dev), n, H, W, cin)  # see _conv_first
        else:
            xin = self._new(n, H, W, cin)
        ops.nchw_to_nhwc(x, xin)
        h = Dual(xin, xin)
        share = shared_halves and self._shareable(n)
        for i, layers in enumerate(self.inp):
            if i == 1 and emb_ready is not None:
                main.wait_event(emb_ready)
                emb_ready = None
            if share and i == 1:
                h = self._run_shared(layers, h, emb_all.buf, kv_all, n_ctx, skip_dst[i])
            else:
                h = self._run(layers, h, emb_all.buf, kv_all, n_ctx, skip_dst[i])
        h = self._run(self.mid, h, emb_all.buf, kv_all, n_ctx, window(0, 0, self.mid_ch))
        for j, layers in enumerate(self.outb):
            if j + 1 < len(self.outb):
                dst = window(j + 1, 0, cat[j + 1].f.c - self.skip_ch[n_in - 2 - j])
            else:
                dst = None
            h = self._run(layers, cat[j], emb_all.buf, kv_all, n_ctx, dst)

        hn = self._new(n, H, W, h.f.c)
        ops.group_norm(h.f, self.out_norm[0], self.out_norm[1], 1e-5, hn, self.stats, silu=True)
        co = m.out_channels
        co_pad = (co + 7) // 8 * 8
        obuf = Act.empty(n, H, W, co_pad, torch.float32, dev)
        ops.conv([ConvIn(hn, 3, 1, 1)], self.out_conv.w, obuf.cols(0, co), bias=self.out_conv.b)
        y = torch.empty((n, co, H, W), dtype=torch.float32, device=dev)
        ops.nhwc_to_nchw(obuf.cols(0, co), y)
        return y


def _dup_first_half(a: Act) -> None:
    """Images n/2 .. n of `a` := images 0 .. n/2 (one device-to-device copy; the statistics buffer follows)."""
    half = a.rows // 2
    v = a.view2d()
    v[half:].copy_(v[:half])
    if a.gp is not None:
        g = a.gp.shape[0] // 2
        a.gp[g:].copy_(a.gp[:g])


class Dual:
    """A residual-stream activation: fp32 master `f` and compute-dtype operand shadow `h` (same object in fp32 mode)."""
    __slots__ = ("f", "h")

    def __init__(self, f: Act, h: Act):
        self.f, self.h = f, h
