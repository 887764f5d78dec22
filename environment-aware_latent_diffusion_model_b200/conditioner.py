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
    def _dropout(self, y: torch.Tensor, p: float = 0.1):
        """nn.Dropout(0.1) of the two MLPs (models.py:320,492) in training mode: the mask is RNG plumbing (torch's
        generator), kept as an inverted-dropout multiplier for the backward.  Returns (y * mask, mask)."""
        if not self.training:
            return y, None
        mask = (torch.rand_like(y) >= p).to(torch.float32).mul_(1.0 / (1.0 - p))
        return y * mask, mask

    def _lstm_mlp(self, p, x: torch.Tensor, save: Optional[dict] = None) -> torch.Tensor:
        """WeatherLSTM.forward: x [B, S, in] -> [B*S, emb]"""
        dev = x.device
        B, S, _ = x.shape
        H = p["w_hh"].shape[1]
        hseq = torch.empty((B, S, H), dtype=torch.float32, device=dev)
        cs = [torch.empty((B, H), dtype=torch.float32, device=dev) for _ in range(S if save is not None else min(S, 2))]
        recs = []
        for t in range(S):
            rec = None
            if t > 0:
                hp = hseq[:, t - 1].contiguous()
                rec = torch.empty((B, 4 * H), dtype=torch.float32, device=dev)
                ops.linear(Act(hp, 1, 1, B), p["w_hh"], Act(rec, 1, 1, B))
            recs.append(rec)
            ops.lstm_cell(x[:, t], p["w_ih"], p["b_ih"], p["b_hh"], hseq[:, t], cs[t % len(cs)], rec=rec,
                          c_prev=cs[(t - 1) % len(cs)] if t > 0 else None)
        hs = hseq.reshape(B * S, H)
        y = torch.empty((B * S, p["fc0"][0].shape[0]), dtype=torch.float32, device=dev)
        ops.linear(Act(hs, 1, 1, B * S), p["fc0"][0], Act(y, 1, 1, B * S), bias=p["fc0"][1], act=L.ACT_RELU)
        yd, mask = self._dropout(y)
        o = torch.empty((B * S, p["fc3"][0].shape[0]), dtype=torch.float32, device=dev)
        ops.linear(Act(yd, 1, 1, B * S), p["fc3"][0], Act(o, 1, 1, B * S), bias=p["fc3"][1])
        if save is not None:
            save.update(x=x, hseq=hseq, cs=cs, recs=recs, y=y, yd=yd, mask=mask)
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

    def _own_named_parameters(self):
        """The conditioner's own parameters (`convs.*` is the frozen first stage, ddpm.py:509-512,535-536)."""
        return [(n, p) for n, p in self.named_parameters() if not n.startswith("convs.")]

    def forward(self, mixed, phase="train", return_intermediates=False):
        """models.py:500-539.  With gradients enabled and trainable parameters the call goes through `_UnetCondFn`, whose
        backward is the hand-written adjoint below (the reference trains the conditioner with the UNet,
        ddpm.py:1409-1415); otherwise it is the plain no-grad forward."""
        own = self._own_named_parameters()
        if torch.is_grad_enabled() and not return_intermediates and any(p.requires_grad for _, p in own):
            return _UnetCondFn.apply(self, mixed, *[p for _, p in own])
        with torch.no_grad():
            return self._forward_impl(mixed, return_intermediates=return_intermediates)

    def _forward_impl(self, mixed, return_intermediates=False, save: Optional[dict] = None):
        if len(mixed) == 4:
            img, flow, weather, time = mixed
        else:
            img, flow, weather, time = mixed[:4]
        have_cond = mixed[-1] is not None     # the negative branch passes (..., None): models.py:516, ddpm.py:886-888
        dev = self.out_layer[1].weight.device
        if dev.type != "cuda":
            raise RuntimeError("ealdm_b200.UnetCond runs on CUDA (sm_100a) only; there is no CPU fallback")
        P = self._pack()
        img = img.squeeze(0).to(dev)
        z = self._encoder_features(img)                                        # [B, 32, 32, mid] NHWC fp32
        B, hw, md = z.n, z.h * z.w, self.mid_dim
        assert z.c == md and hw == 32 * 32, "out_layer is built for a 32 x 32 x mid_dim encoder output (models.py:489)"
        inter = {}
        feat = z
        S = save if save is not None else None
        if S is not None:
            S.update(P=P, z=z, have_cond=have_cond, B=B)
        if have_cond:
            weather = weather.squeeze(0).float().to(dev)
            flow = flow.squeeze(0).float().to(dev)
            time = time.squeeze(0).float().to(dev).reshape(B, 1).contiguous()
            t_sty = torch.empty((B, self.emb_dim), dtype=torch.float32, device=dev)
            four = torch.empty((B, 2 * self._freqs.numel()), dtype=torch.float32, device=dev) if S is not None else None
            ops.fourier_style(time, self._freqs.to(dev), bool(self.cond_args.get("include_lin", False)),
                              float(self.cond_args.get("lin_lr", 0.0)), P["scale_w"],
                              self.scaled_styles.c_to_scales.weight_gain, t_sty, features=four)
            sf, sw = ({} if S is not None else None), ({} if S is not None else None)
            f_sty = self._lstm_mlp(P["f_mlp"], flow, sf)
            w_sty = self._lstm_mlp(P["w_mlp"], weather, sw)
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
            if S is not None:   # the backward re-derives the statistics the forward normalised with
                S.update(four=four, sf=sf, sw=sw, styles=(w_sty, f_sty, t_sty), cat=cat, y0=y0, y1=y1, bn_train=bn_train,
                         bn_mean=bn.running_mean.clone(), bn_var=bn.running_var.clone())
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
        h1d, mask1 = self._dropout(h1)       # models.py:492
        ctx = torch.empty((B * md, self.out_dim), dtype=torch.float32, device=dev)
        ops.linear(Act(h1d, 1, 1, B * md), P["out4"][0], Act(ctx, 1, 1, B * md), bias=P["out4"][1])
        if S is not None:
            S.update(x2=x2, h1=h1, h1d=h1d, mask1=mask1)
        ctx = ctx.view(B, md, self.out_dim)
        if return_intermediates:
            inter["mixed"] = rows
            return ctx, inter
        return ctx

    # ---- adjoint -----------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def _backward_impl(self, S: dict, dctx: torch.Tensor):
        """Gradients of every own parameter (dict name -> tensor) for d(loss)/d(context) = dctx [B, mid, out_dim].
        Linear / conv data gradients run on the same conv kernels with transposed (and tap-flipped) weights, weight
        gradients on ealdm_conv_wgrad (fp32), bias gradients on ealdm_colsum; AdaIN / BatchNorm+ReLU / LSTM-cell / ReLU
        adjoints are the four kernels of csrc/cond.cu.  The frozen encoder output z receives no gradient."""
        from .train import _pack_dgrad
        P, z, B = S["P"], S["z"], S["B"]
        dev, md, hw = dctx.device, self.mid_dim, z.h * z.w
        f32 = torch.float32
        ws = ops.Workspace(dev)
        G = {n: torch.zeros_like(p, dtype=f32) for n, p in self._own_named_parameters()}
        act2 = lambda t2: Act(t2, 1, 1, t2.shape[0])  # noqa: E731

        linear_bwd = lambda *a, **k: self._linear_bwd(G, ws, *a, **k)  # noqa: E731

        d2 = dctx.reshape(B * md, self.out_dim).float().contiguous()
        dh1 = linear_bwd(S["h1d"], d2, "out_layer.4.weight", "out_layer.4.bias")
        dpre1 = ops.relu_bwd(S["h1"], dh1, S["mask1"])
        dx2 = linear_bwd(S["x2"], dpre1, "out_layer.1.weight", "out_layer.1.bias", need_dx=S["have_cond"])
        if not S["have_cond"]:
            return G
        dfeat = Act.empty(z.n, z.h, z.w, md, f32, dev)
        ops.nchw_to_nhwc(dx2.view(B, md, z.h, z.w), dfeat)
        # conv_cat[3]: 3x3, mid -> mid (+ residual z, frozen)
        w3 = self.conv_cat[3].weight
        ops.conv_wgrad(S["y1"], dfeat, G["conv_cat.3.weight"].view(w3.shape[0], -1), ws, ksize=3, stride=1, pad=1)
        ops.colsum(dfeat, G["conv_cat.3.bias"], ws)
        dy1 = Act.empty(z.n, z.h, z.w, md, f32, dev)
        ops.conv([ConvIn(dfeat, 3, 1, 1)], _pack_dgrad(w3, f32), dy1)
        # BatchNorm2d + ReLU
        bn = self.conv_cat[1]
        dy0 = Act.empty(z.n, z.h, z.w, md, f32, dev)
        ops.batch_norm_relu_bwd(S["y0"], S["y1"], dy1, dy0, bn.weight.detach().float(), S["bn_mean"], S["bn_var"],
                                training=S["bn_train"], eps=bn.eps, relu=True, dgamma=G["conv_cat.1.weight"],
                                dbeta=G["conv_cat.1.bias"])
        # conv_cat[0]: 3x3, 4 mid -> mid over [z | AdaIN_w | AdaIN_f | AdaIN_t]
        w0 = self.conv_cat[0].weight
        ops.conv_wgrad(S["cat"], dy0, G["conv_cat.0.weight"].view(w0.shape[0], -1), ws, ksize=3, stride=1, pad=1)
        ops.colsum(dy0, G["conv_cat.0.bias"], ws)
        dcat = Act.empty(z.n, z.h, z.w, 4 * md, f32, dev)
        ops.conv([ConvIn(dy0, 3, 1, 1)], _pack_dgrad(w0, f32), dcat)
        dsty = []
        for k, (name, sty) in enumerate(zip(("wadain", "fadain", "tadain"), S["styles"])):
            daff = torch.empty((B, 2 * md), dtype=f32, device=dev)
            ops.adain_bwd(z, dcat.cols((k + 1) * md, md), daff, eps=1e-5)
            dsty.append(linear_bwd(sty, daff, f"{name}.linear.weight", f"{name}.linear.bias"))
        dw_sty, df_sty, dt_sty = dsty
        # CondScale: t_sty = fourier(time) (W gain)^T
        gw = G["scaled_styles.c_to_scales.weight"]
        tmp = torch.zeros_like(gw)
        ops.linear_wgrad(act2(S["four"]), act2(dt_sty), tmp, ws)
        gw.add_(tmp, alpha=float(self.scaled_styles.c_to_scales.weight_gain))
        for prefix, sv, do in (("f_mlp", S["sf"], df_sty), ("w_mlp", S["sw"], dw_sty)):
            self._lstm_mlp_bwd(prefix, P[prefix], sv, do, G, ws, linear_bwd)
        return G

    def _linear_bwd(self, G, ws, x2d, dy2d, wname, bname, need_dx=True):
        """y = x W^T + b: accumulates dW into G[wname] and db into G[bname]; returns dx = dy W."""
        from .train import _pack_dgrad
        w = self.get_parameter(wname)
        act2 = lambda t2: Act(t2, 1, 1, t2.shape[0])  # noqa: E731
        ops.linear_wgrad(act2(x2d), act2(dy2d), G[wname].view(w.shape[0], -1), ws)
        if bname is not None:
            ops.colsum(act2(dy2d), G[bname], ws)
        if not need_dx:
            return None
        dx = torch.empty((dy2d.shape[0], w.shape[1]), dtype=torch.float32, device=dy2d.device)
        ops.linear(act2(dy2d), _pack_dgrad(w, torch.float32), act2(dx))
        return dx

    def _lstm_mlp_bwd(self, prefix, p, sv, do, G, ws, linear_bwd):
        """WeatherLSTM adjoint: Linear - (Dropout) - ReLU - Linear, then back-propagation through time."""
        from .train import _pack_dgrad
        dev, f32 = do.device, torch.float32
        x, hseq, cs, recs = sv["x"], sv["hseq"], sv["cs"], sv["recs"]
        B, Sq, H = hseq.shape
        dyd = linear_bwd(sv["yd"], do, f"{prefix}.fc.3.weight", f"{prefix}.fc.3.bias")
        dpre = ops.relu_bwd(sv["y"], dyd, sv["mask"])
        dhs = linear_bwd(hseq.reshape(B * Sq, H), dpre, f"{prefix}.fc.0.weight", f"{prefix}.fc.0.bias").view(B, Sq, H)
        w_hh_t = _pack_dgrad(p["w_hh"], f32) if Sq > 1 else None
        dh_rec, dc_next = None, None
        act2 = lambda t2: Act(t2, 1, 1, t2.shape[0])  # noqa: E731
        for t in range(Sq - 1, -1, -1):
            dh = dhs[:, t] if dh_rec is None else (dhs[:, t] + dh_rec)
            dgates = torch.empty((B, 4 * H), dtype=f32, device=dev)
            dc_prev = torch.empty((B, H), dtype=f32, device=dev) if t > 0 else None
            ops.lstm_cell_bwd(x[:, t], p["w_ih"], p["b_ih"], p["b_hh"], dh, dgates, rec=recs[t],
                              c_prev=cs[t - 1] if t > 0 else None, dc_next=dc_next, dc_prev=dc_prev)
            xt = x[:, t].contiguous()
            ops.linear_wgrad(act2(xt), act2(dgates), G[f"{prefix}.lstm.weight_ih_l0"], ws)
            ops.colsum(act2(dgates), G[f"{prefix}.lstm.bias_ih_l0"], ws)
            ops.colsum(act2(dgates), G[f"{prefix}.lstm.bias_hh_l0"], ws)
            if t > 0:
                ops.linear_wgrad(act2(hseq[:, t - 1].contiguous()), act2(dgates), G[f"{prefix}.lstm.weight_hh_l0"], ws)
                dh_rec = torch.empty((B, H), dtype=f32, device=dev)
                ops.linear(act2(dgates), w_hh_t, act2(dh_rec))
            dc_next = dc_prev


class _UnetCondFn(torch.autograd.Function):
    """UnetCond.forward under torch.autograd: the parameters are inputs so that `loss.backward()` (the UNet's own
    hand-written backward hands back d(loss)/d(context)) reaches them."""

    @staticmethod
    def forward(ctx, module, mixed, *params):
        saved = {}
        out = module._forward_impl(mixed, save=saved)
        ctx.module, ctx.saved = module, saved
        ctx.names = [n for n, _ in module._own_named_parameters()]
        return out

    @staticmethod
    def backward(ctx, dctx):
        G = ctx.module._backward_impl(ctx.saved, dctx.contiguous())
        ctx.saved = None
        needs = ctx.needs_input_grad[2:]
        return (None, None) + tuple(G[n] if need else None for n, need in zip(ctx.names, needs))
