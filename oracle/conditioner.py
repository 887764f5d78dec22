"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- CPU fp32 restatement of the EALDM conditioner `UnetCond`
(SURVEY.md section 8f rank 3), the module that turns (frames, optical-flow magnitude, weather vector, time stamp)
into the [T, 4, 512] cross-attention context of the stdiff UNet.

Follows STDiff/models.py of the reference:
  UnetCond.__init__ / forward      :411-539   (the `convs` sub-module is REPLACED by the first stage after construction,
                                               ldm/models/diffusion/ddpm.py:535-536, and only `convs.encoder` is called)
  ConditioningTransform.forward    :203-236   (fourier features of the time stamp, explicit linear term)
  CondScale.forward                :298-309   (one bias-free FullyConnectedLayer, weight * lr_multiplier / sqrt(in))
  FullyConnectedLayer.forward      :262-274
  WeatherLSTM.forward              :323-336   (nn.LSTM batch-first over [B, len_seq, in], len_seq = 1 in the shipped data; then a 2-layer MLP)
  AdaIN.forward                    :369-377   (InstanceNorm2d, x * (1 + gamma) + beta from a Linear of the style)
Dropout(0.1) layers are identities in eval mode; BatchNorm2d uses running statistics in eval mode and the statistics
of the batch (biased variance over T*H*W) in training mode -- both are restated.

Pinned by tests/golden/conditioner.pt, produced by the reference's own UnetCond (oracle/gen_golden_cond.py).
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

# cond_args of configs/latent-diffusion/stdiff_cin-ldm-vq-f8.yaml:65-81
COND_ARGS = {"type": "fourier", "dequant": "gauss", "noise": 0, "noise_f_int": [None],
             "noise_f": [263.246328125, 7.791666666666667, 0], "dims": 6, "lr": 1, "lin_lr": 0.01,
             "f_manual": [1.839835728952772, 672], "include_lin": True}
DEFAULTS = {"mid_dim": 4, "emb_dim": 128, "out_dim": 512, "w_dim": 16, "f_dim": 1, "t_dim": 6, "hidden_dim": 1024,
            "num_layers": 1, "num_ws": 1}


def param_shapes(cfg: dict = DEFAULTS) -> List[Tuple[str, Tuple[int, ...]]]:
    """state_dict inventory of UnetCond WITHOUT `convs.*` (the first stage plugged in by LatentDiffusion), in the
    reference's registration order (models.py:457-497), BatchNorm buffers included."""
    md, ed, od, hd = cfg["mid_dim"], cfg["emb_dim"], cfg["out_dim"], cfg["hidden_dim"]
    out: List[Tuple[str, Tuple[int, ...]]] = []

    def lstm(prefix, in_dim):
        out.extend([(f"{prefix}.lstm.weight_ih_l0", (4 * hd, in_dim)), (f"{prefix}.lstm.weight_hh_l0", (4 * hd, hd)),
                    (f"{prefix}.lstm.bias_ih_l0", (4 * hd,)), (f"{prefix}.lstm.bias_hh_l0", (4 * hd,)),
                    (f"{prefix}.fc.0.weight", (ed, hd)), (f"{prefix}.fc.0.bias", (ed,)),
                    (f"{prefix}.fc.3.weight", (ed, ed)), (f"{prefix}.fc.3.bias", (ed,))])

    def adain(prefix):
        out.extend([(f"{prefix}.linear.weight", (2 * md, ed)), (f"{prefix}.linear.bias", (2 * md,))])

    lstm("w_mlp", cfg["w_dim"])
    adain("wadain")
    lstm("f_mlp", cfg["f_dim"])
    adain("fadain")
    out.append(("scaled_styles.c_to_scales.weight", (ed, cfg["t_dim"])))
    adain("tadain")
    out.extend([("conv_cat.0.weight", (md, 4 * md, 3, 3)), ("conv_cat.0.bias", (md,)),
                ("conv_cat.1.weight", (md,)), ("conv_cat.1.bias", (md,)),
                ("conv_cat.1.running_mean", (md,)), ("conv_cat.1.running_var", (md,)),
                ("conv_cat.1.num_batches_tracked", ()),
                ("conv_cat.3.weight", (md, md, 3, 3)), ("conv_cat.3.bias", (md,)),
                ("out_layer.1.weight", (md * 32 * 32, 32 * 32)), ("out_layer.1.bias", (md * 32 * 32,)),
                ("out_layer.4.weight", (od, md * 32 * 32)), ("out_layer.4.bias", (od,))])
    return out


def synthetic_state_dict(cfg: dict = DEFAULTS, seed: int = 9) -> SD:
    """Deterministic weights, reference-independent: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (nn.Linear / nn.LSTM default
    scale) for matrices and biases, BatchNorm gain ~ U(0.5, 1.5), running statistics away from (0, 1) so that the
    eval-mode path is exercised, and the CondScale matrix at O(1) (its reference init 1e-6 * randn + e_0 would make
    the time style a constant)."""
    sd: SD = {}
    by_name = dict(param_shapes(cfg))
    for idx, (name, shape) in enumerate(param_shapes(cfg)):
        g = torch.Generator().manual_seed(seed * 100003 + idx)
        if name.endswith("num_batches_tracked"):
            t = torch.tensor(3, dtype=torch.long)
        elif name.endswith("running_mean"):
            t = torch.randn(shape, generator=g) * 0.3
        elif name.endswith("running_var"):
            t = torch.rand(shape, generator=g) + 0.5
        elif name == "conv_cat.1.weight":
            t = torch.rand(shape, generator=g) + 0.5
        elif name == "conv_cat.1.bias":
            t = torch.randn(shape, generator=g) * 0.2
        elif ".lstm." in name:
            bound = 1.0 / math.sqrt(cfg["hidden_dim"])
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif name == "scaled_styles.c_to_scales.weight":
            t = torch.randn(shape, generator=g)
        else:
            wshape = by_name[name[:-5] + ".weight"] if name.endswith(".bias") else shape
            bound = 1.0 / math.sqrt(math.prod(wshape[1:]))
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        sd[name] = t
    return sd


def fourier_features(time: torch.Tensor, cond_args: dict = COND_ARGS) -> torch.Tensor:
    """ConditioningTransform.forward(time, broadcast=True) followed by cs.unbind(dim=1)[0] (models.py:203-236,511-513):
    time [T, 1] -> [T, 2 * #freq] = [cos0, sin0, cos1, sin1, ...] over the ascending frequencies (-1 first when the
    explicit linear term is on: its pair is overwritten by (1, lin_lr * t))."""
    freqs = list(cond_args["f_manual"])
    if cond_args.get("include_lin", False):
        freqs = [-1.0] + freqs
    fr = torch.from_numpy(np.sort(freqs).astype(np.float32))
    cos = torch.cos(2 * np.pi * fr * time)
    sin = torch.sin(2 * np.pi * fr * time)
    if cond_args.get("include_lin", False):
        cos[:, 0] = 1
        sin[:, 0] = cond_args["lin_lr"] * time[:, 0]
    return torch.stack((cos, sin), dim=-1).view(time.shape[0], -1)


def cond_scale(sd: SD, c: torch.Tensor, cond_args: dict = COND_ARGS) -> torch.Tensor:
    """CondScale.forward(c=c): FullyConnectedLayer without bias, runtime weight gain lr / sqrt(in_features)."""
    w = sd["scaled_styles.c_to_scales.weight"]
    gain = cond_args["lr"] / math.sqrt(w.shape[1])
    return c.matmul((w * gain).t())


def lstm_mlp(sd: SD, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """WeatherLSTM.forward: x [B, S, in] batch-first, zero initial state, gate order i, f, g, o (torch.nn.LSTM); every
    step's hidden state goes through Linear - ReLU - (Dropout) - Linear -> [B * S, emb].  The shipped data format has
    S = 1 (dataset_wlbl.py:567-570: one frame per item, `len_seq: 1` in the yaml), i.e. ONE cell step per frame, and the
    AdaIN that consumes the result only broadcasts for S = 1; longer sequences are restated anyway."""
    w_ih, w_hh = sd[f"{prefix}.lstm.weight_ih_l0"], sd[f"{prefix}.lstm.weight_hh_l0"]
    b = sd[f"{prefix}.lstm.bias_ih_l0"] + sd[f"{prefix}.lstm.bias_hh_l0"]
    hd = w_hh.shape[1]
    B, S = x.shape[0], x.shape[1]
    h = torch.zeros(B, hd)
    c = torch.zeros(B, hd)
    hs = []
    for t in range(S):
        gates = x[:, t] @ w_ih.t() + h @ w_hh.t() + b
        i, f, g, o = gates.split(hd, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
        hs.append(h)
    hseq = torch.stack(hs, dim=1).reshape(B * S, hd)
    y = F.relu(F.linear(hseq, sd[f"{prefix}.fc.0.weight"], sd[f"{prefix}.fc.0.bias"]))
    return F.linear(y, sd[f"{prefix}.fc.3.weight"], sd[f"{prefix}.fc.3.bias"])


def adain(sd: SD, prefix: str, x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """AdaIN.forward: instance norm (biased variance over H*W, eps 1e-5, no affine) then x * (1 + gamma) + beta."""
    xn = F.instance_norm(x, eps=1e-5)
    h = F.linear(w, sd[f"{prefix}.linear.weight"], sd[f"{prefix}.linear.bias"])
    gamma, beta = h.chunk(2, 1)
    return xn * (1 + gamma[:, :, None, None]) + beta[:, :, None, None]


def unet_cond_forward(sd: SD, z: torch.Tensor, flow: torch.Tensor, weather: torch.Tensor, time: torch.Tensor,
                      bn_training: bool = False, cond_args: dict = COND_ARGS) -> Dict[str, torch.Tensor]:
    """UnetCond.forward after `img = self.convs.encoder(img)` (models.py:515-539): z [T, 4, 32, 32] is the first-stage
    encoder output, flow [T, 1, f_dim], weather [T, 1, w_dim], time [T, 1] (the DataLoader batch of dataset_wlbl.py).  Returns the context [T, 4, out_dim] and the
    intermediates the GPU tests compare one by one."""
    c = fourier_features(time, cond_args)
    t_sty = cond_scale(sd, c, cond_args)
    f_sty = lstm_mlp(sd, "f_mlp", flow)
    w_sty = lstm_mlp(sd, "w_mlp", weather)
    ws, fs, ts = adain(sd, "wadain", z, w_sty), adain(sd, "fadain", z, f_sty), adain(sd, "tadain", z, t_sty)
    cat = torch.cat((z, ws, fs, ts), dim=1)
    y = F.conv2d(cat, sd["conv_cat.0.weight"], sd["conv_cat.0.bias"], padding=1)
    y = F.batch_norm(y, None if bn_training else sd["conv_cat.1.running_mean"],
                     None if bn_training else sd["conv_cat.1.running_var"], sd["conv_cat.1.weight"],
                     sd["conv_cat.1.bias"], training=bn_training, eps=1e-5)
    y = F.conv2d(F.relu(y), sd["conv_cat.3.weight"], sd["conv_cat.3.bias"], padding=1)
    img = y + z
    h = F.relu(F.linear(img.flatten(2), sd["out_layer.1.weight"], sd["out_layer.1.bias"]))
    ctx = F.linear(h, sd["out_layer.4.weight"], sd["out_layer.4.bias"])
    return {"context": ctx, "time_style": t_sty, "flow_style": f_sty, "weather_style": w_sty, "mixed": img}
