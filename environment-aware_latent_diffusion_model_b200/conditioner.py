"""The EALDM conditioner `UnetCond` (reference: STDiff/models.py:411-539) on libealdm_b200: the module that turns
(frames, flow magnitude, weather vector, time stamp) into the [B, 4, 512] cross-attention context of the stdiff UNet
(`cond_stage_config.target` of configs/latent-diffusion/stdiff_cin-ldm-vq-f8.yaml:61-81).

Same constructor keywords, same parameter / buffer names (everything except `convs.*`: the reference builds a torchvision
ResNet-50 there and `LatentDiffusion.instantiate_cond_stage` immediately replaces it by the first stage,
ddpm.py:535-536 -- here `convs` starts as None and must be assigned the same way), same `forward(mixed, phase)` contract.

Data path (all fp32 except the encoder, which runs in the first stage's compute dtype):
  first-stage encoder (autoencoder.py: AutoencoderEngine.encoder_features, tcgen05 convs)        -> z [B, 32, 32, 4] NHWC
  ealdm_fourier_style: ConditioningTransform + CondScale of the time stamp (models.py:203-236, 298-309)  -> [B, 128]
  ealdm_lstm_cell (+ ealdm_conv as W_hh h for sequences longer than 1) + 2 linear layers (models.py:312-336) x 2
  ealdm_adain x 3 into column windows of one [B*1024, 16] concat buffer (no torch.cat copy)    (models.py:362-377)
  conv3x3 16->4, ealdm_batch_norm_relu, conv3x3 4->4 + residual z                              (models.py:480-483, 527-529)
  NHWC -> [B*4, 1024] rows, Linear 1024->4096 + ReLU, Linear 4096->512                          (models.py:485-494)
No CPU fallback: inputs are moved to the module's CUDA device, and a missing library raises.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib as L
from . import ops
from .ops import Act, ConvIn
from .packing import pack_conv_weight


class _Args(dict):
    __getattr__ = dict.get


class WeatherLSTM(nn.Module):
    """models.py:312-336 (parameter container; `nn.LSTM` only holds weight_ih_l0 / weight_hh_l0 / bias_*_l0)"""

    def __init__(self, input_size, hidden_size, num_layers, output_size):
        super().__init__()
        if num_layers != 1:
            raise NotImplementedError("UnetCond uses single-layer LSTMs (models.py:425)")
        self.hidden_size = hidden_size
        self.lstm = nn.LSTM(input_size, hidden_size, num_layers, batch_first=True)
        self.fc = nn.Sequential(nn.Linear(hidden_size, output_size), nn.ReLU(), nn.Dropout(0.1),
                                nn.Linear(output_size, output_size))


class AdaIN(nn.Module):
    """models.py:362-377 (parameter container)"""

    def __init__(self, in_dim, w_dim):
        super().__init__()
        self.in_dim = in_dim
        self.linear = nn.Linear(w_dim, in_dim * 2)


class _FC(nn.Module):
    """FullyConnectedLayer without bias (models.py:239-274): runtime gain lr_multiplier / sqrt(in_features)"""

    def __init__(self, in_features, out_features, lr_multiplier=1.0):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_features, in_features) / lr_multiplier)
        self.weight_gain = lr_multiplier / math.sqrt(in_features)


class CondScale(nn.Module):
    """models.py:283-309 with w_dim=None, cond_args.type == 'fourier'"""

    def __init__(self, c_dim, channels, cond_args):
        super().__init__()
        self.c_to_scales = _FC(c_dim, channels, lr_multiplier=cond_args.get("lr", 1))
        with torch.no_grad():
            self.c_to_scales.weight.mul_(1e-6)
            self.c_to_scales.weight[:, 0] += 1


class UnetCond(nn.Module):
    def __init__(self, dim=64, init_dim=None, mid_dim=4, emb_dim=128, out_dim=512, dim_mults=(1, 2, 4, 8), channels=3,
                 resnet_block_groups=8, w_dim=16, f_dim=1, t_dim=6, hidden_dim=1024, num_layers=1, num_ws=1,
                 cond_args=None, device=None):
        super().__init__()
        cond_args = _Args(cond_args or {})
        if cond_args.get("type") != "fourier":
            raise NotImplementedError("UnetCond: only the 'fourier' time conditioning of the shipped config is built")
        self.cond_args, self.mid_dim, self.emb_dim, self.out_dim, self.num_ws = cond_args, mid_dim, emb_dim, out_dim, num_ws
        self.convs: Optional[nn.Module] = None          # the first stage, assigned by LatentDiffusion (ddpm.py:535-536)
        self.w_mlp = WeatherLSTM(w_dim, hidden_dim, num_layers, emb_dim)
        self.wadain = AdaIN(mid_dim, emb_dim)
        self.f_mlp = WeatherLSTM(f_dim, hidden_dim, num_layers, emb_dim)
        self.fadain = AdaIN(mid_dim, emb_dim)
        self.scaled_styles = CondScale(t_dim, emb_dim, cond_args)
        self.tadain = AdaIN(mid_dim, emb_dim)
        self.conv_cat = nn.Sequential(nn.Conv2d(4 * mid_dim, mid_dim, 3, 1, 1), nn.BatchNorm2d(mid_dim), nn.ReLU(),
                                      nn.Conv2d(mid_dim, mid_dim, 3, 1, 1))
        self.out_layer = nn.Sequential(nn.Flatten(2), nn.Linear(32 * 32, mid_dim * 32 * 32), nn.ReLU(), nn.Dropout(0.1),
                                       nn.Linear(mid_dim * 32 * 32, out_dim))
        for name, module in self.named_children():       # models.py:496-505
            if name != "convs":
                module.apply(self._init_weights)
        freqs = list(cond_args.get("f_manual") or [])
        if cond_args.get("include_lin", False):
            freqs = [-1.0] + freqs
        self.register_buffer("_freqs", torch.from_numpy(np.sort(freqs).astype(np.float32)), persistent=False)
        assert 2 * len(freqs) == t_dim, "t_dim must equal cond_args.dims = 2 * #frequencies"
        self._packed = None
        self._warned = False
        self.register_load_state_dict_post_hook(lambda m, keys: m.invalidate_packed())
        if device is not None and str(device) != "cpu":
            self.to(device)

    @staticmethod
    def _init_weights(module):
        if isinstance(module, (nn.Linear, nn.Conv2d)):
            nn.init.kaiming_normal_(module.weight.data, mode="fan_out", nonlinearity="relu")
            if isinstance(module, nn.Linear) and module.bias is not None:
                module.bias.data.zero_()

    def invalidate_packed(self):
        self._packed = None

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    # ---- kernel-layout copies of the parameters (made once; call invalidate_packed() after an optimizer step) --------
    def _pack(self):
        own = [p for n, p in self.named_parameters() if not n.startswith("convs.")]
        fp = sum(p._version for p in own)
        if self._packed is not None and fp == self._packed_fp:
            return self._packed
        self._packed_fp = fp
        f32 = lambda t: t.detach().float().contiguous()  # noqa: E731
        P = {}
        for name in ("w_mlp", "f_mlp"):
            m = getattr(self, name)
            P[name] = {"w_ih": f32(m.lstm.weight_ih_l0), "w_hh": f32(m.lstm.weight_hh_l0), "b_ih": f32(m.lstm.bias_ih_l0),
                       "b_hh": f32(m.lstm.bias_hh_l0), "fc0": (f32(m.fc[0].weight), f32(m.fc[0].bias)),
                       "fc3": (f32(m.fc[3].weight), f32(m.fc[3].bias))}
        for name in ("wadain", "fadain", "tadain"):
            m = getattr(self, name)
            P[name] = (f32(m.linear.weight), f32(m.linear.bias))
        P["scale_w"] = f32(self.scaled_styles.c_to_scales.weight)
        P["conv0"] = (pack_conv_weight(self.conv_cat[0].weight.detach(), torch.float32), f32(self.conv_cat[0].bias))
        P["conv3"] = (pack_conv_weight(self.conv_cat[3].weight.detach(), torch.float32), f32(self.conv_cat[3].bias))
        P["out1"] = (f32(self.out_layer[1].weight), f32(self.out_layer[1].bias))
        P["out4"] = (f32(self.out_layer[4].weight), f32(self.out_layer[4].bias))
        self._packed = P
        return P

    # ---- pieces ------------------------------------------------------------------------------------------------------
    def _lstm_mlp(self, p, x: torch.Tensor) -> torch.Tensor:
        """WeatherLSTM.forward: x [B, S, in] -> [B*S, emb]"""
        dev = x.device
        B, S, _ = x.shape
        H = p["w_hh"].shape[1]
        hseq = torch.empty((B, S, H), dtype=torch.float32, device=dev)
        c = [torch.empty((B, H), dtype=torch.float32, device=dev) for _ in range(2)]
        rec = None
        for t in range(S):
            if t > 0:
                hp = hseq[:, t - 1].contiguous()
                rec = torch.empty((B, 4 * H), dtype=torch.float32, device=dev)
                ops.linear(Act(hp, 1, 1, B), p["w_hh"], Act(rec, 1, 1, B))
            ops.lstm_cell(x[:, t], p["w_ih"], p["b_ih"], p["b_hh"], hseq[:, t], c[t & 1], rec=rec,
                          c_prev=c[(t - 1) & 1] if t > 0 else None)
        hs = hseq.reshape(B * S, H)
        y = torch.empty((B * S, p["fc0"][0].shape[0]), dtype=torch.float32, device=dev)
        ops.linear(Act(hs, 1, 1, B * S), p["fc0"][0], Act(y, 1, 1, B * S), bias=p["fc0"][1], act=L.ACT_RELU)
        if self.training:
            y = F.dropout(y, 0.1)       # RNG plumbing only (models.py:320); the mask comes from torch's generator
        o = torch.empty((B * S, p["fc3"][0].shape[0]), dtype=torch.float32, device=dev)
        ops.linear(Act(y, 1, 1, B * S), p["fc3"][0], Act(o, 1, 1, B * S), bias=p["fc3"][1])
        return o

    def _encoder_features(self, img: torch.Tensor) -> Act:
        if self.convs is None:
            raise RuntimeError("UnetCond.convs is None: assign the first stage (ldm/models/diffusion/ddpm.py:535-536)")
        eng = getattr(self.convs, "_eng", None)
        if eng is None:
            raise TypeError("UnetCond.convs must be an ealdm_b200 first stage (AutoencoderKL / VQModelInterface)")
        z = eng(img).encoder_features(img.float().contiguous())
        if z.dtype != torch.float32:
            z32 = Act.empty(z.n, z.h, z.w, z.c, torch.float32, img.device)
            ops.copy2d(z, z32)
            z = z32
        return z

    @torch.no_grad()
    def forward(self, mixed, phase="train", return_intermediates=False):
        if len(mixed) == 4:
            img, flow, weather, time = mixed
            have_cond = mixed[-1] is not None
        else:
            img, flow, weather, time = mixed[:4]
            have_cond = mixed[-1] is not None     # the negative branch passes (..., None): models.py:516, ddpm.py:1322-1324
        dev = self.out_layer[1].weight.device
        if dev.type != "cuda":
            raise RuntimeError("ealdm_b200.UnetCond runs on CUDA (sm_100a) only; there is no CPU fallback")
        if self.training and not self._warned and any(p.requires_grad for p in self.out_layer.parameters()):
            # be loud: a `cond_stage_trainable: true` run would otherwise train the UNet only without saying so
            import warnings
            warnings.warn("ealdm_b200.UnetCond: only the forward pass is built; the conditioner's parameters receive no "
                          "gradients (DESIGN.md section 7). Freeze it (cond_stage_trainable: false) or train it with the "
                          "reference module.", RuntimeWarning, stacklevel=2)
            self._warned = True
        P = self._pack()
        img = img.squeeze(0).to(dev)
        z = self._encoder_features(img)                                        # [B, 32, 32, mid] NHWC fp32
        B, hw, md = z.n, z.h * z.w, self.mid_dim
        assert z.c == md and hw == 32 * 32, "out_layer is built for a 32 x 32 x mid_dim encoder output (models.py:489)"
        inter = {}
        feat = z
        if have_cond:
            weather = weather.squeeze(0).float().to(dev)
            flow = flow.squeeze(0).float().to(dev)
            time = time.squeeze(0).float().to(dev).reshape(B, 1).contiguous()
            t_sty = torch.empty((B, self.emb_dim), dtype=torch.float32, device=dev)
            ops.fourier_style(time, self._freqs.to(dev), bool(self.cond_args.get("include_lin", False)),
                              float(self.cond_args.get("lin_lr", 0.0)), P["scale_w"],
                              self.scaled_styles.c_to_scales.weight_gain, t_sty)
            f_sty = self._lstm_mlp(P["f_mlp"], flow)
            w_sty = self._lstm_mlp(P["w_mlp"], weather)
            cat = Act(torch.empty((B * hw, 4 * md), dtype=torch.float32, device=dev), z.n, z.h, z.w)
            ops.copy2d(z, cat.cols(0, md))
            for k, (name, sty) in enumerate((("wadain", w_sty), ("fadain", f_sty), ("tadain", t_sty))):
                aff = torch.empty((B, 2 * md), dtype=torch.float32, device=dev)
                ops.linear(Act(sty, 1, 1, B), P[name][0], Act(aff, 1, 1, B), bias=P[name][1])
                ops.adain(z, aff, cat.cols((k + 1) * md, md), eps=1e-5)
            y0 = Act.empty(z.n, z.h, z.w, md, torch.float32, dev)
            ops.conv([ConvIn(cat, 3, 1, 1)], P["conv0"][0], y0, bias=P["conv0"][1])
            bn = self.conv_cat[1]
            bn_train = bn.training
            stats = torch.empty((2, md), dtype=torch.float32, device=dev) if bn_train else None
            y1 = Act.empty(z.n, z.h, z.w, md, torch.float32, dev)
            ops.batch_norm_relu(y0, bn.weight.detach().float(), bn.bias.detach().float(), bn.running_mean, bn.running_var,
                                y1, training=bn_train, eps=bn.eps, relu=True, batch_stats=stats)
            if bn_train and bn.track_running_stats:     # nn.BatchNorm2d side effect: momentum update, unbiased variance
                mom = bn.momentum if bn.momentum is not None else 0.1
                nrow = B * hw
                bn.running_mean.mul_(1 - mom).add_(stats[0], alpha=mom)
                bn.running_var.mul_(1 - mom).add_(stats[1] * (nrow / max(nrow - 1, 1)), alpha=mom)
                bn.num_batches_tracked += 1
            feat = Act.empty(z.n, z.h, z.w, md, torch.float32, dev)
            ops.conv([ConvIn(y1, 3, 1, 1)], P["conv3"][0], feat, bias=P["conv3"][1], residual=z)
            inter.update(time_style=t_sty, flow_style=f_sty, weather_style=w_sty)
        # out_layer: Flatten(2) of the NCHW map = rows (frame, channel), 1024 features each
        rows = torch.empty((B, md, z.h, z.w), dtype=torch.float32, device=dev)
        ops.nhwc_to_nchw(feat, rows)
        x2 = rows.view(B * md, hw)
        h1 = torch.empty((B * md, P["out1"][0].shape[0]), dtype=torch.float32, device=dev)
        ops.linear(Act(x2, 1, 1, B * md), P["out1"][0], Act(h1, 1, 1, B * md), bias=P["out1"][1], act=L.ACT_RELU)
        if self.training:
            h1 = F.dropout(h1, 0.1)      # models.py:492, RNG plumbing as above
        ctx = torch.empty((B * md, self.out_dim), dtype=torch.float32, device=dev)
        ops.linear(Act(h1, 1, 1, B * md), P["out4"][0], Act(ctx, 1, 1, B * md), bias=P["out4"][1])
        ctx = ctx.view(B, md, self.out_dim)
        if return_intermediates:
            inter["mixed"] = rows
            return ctx, inter
        return ctx
