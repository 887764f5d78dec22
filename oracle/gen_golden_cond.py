"""TEST INFRASTRUCTURE ONLY -- golden vectors of the EALDM conditioner, made by the
reference's own `STDiff.models.UnetCond` (STDiff/models.py:411-539) with the reference's `VQModelInterface` plugged in
as `convs` exactly as `LatentDiffusion.instantiate_cond_stage` does (ldm/models/diffusion/ddpm.py:535-536).

Shims on top of oracle/gen_golden.py's: `torch.cuda.current_device()` (evaluated in default arguments at import time,
models.py:246,428) returns "cpu"; `torchvision.models.resnet50(pretrained=True)` (models.py:455: a download, and the
module it builds is thrown away when `convs` is replaced) builds the un-pretrained network.

-> tests/golden/conditioner.pt: the inputs' seeds and small tensors, the encoder output, the three styles, the mixed
feature map and the context [T, 4, 512], in eval mode and with BatchNorm in training mode (batch statistics).
Run in the build container (needs /root/reference):  python oracle/gen_golden_cond.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

T = 5  # frames


class _Args(dict):
    __getattr__ = dict.get


def make_inputs():
    g = lambda s: torch.Generator().manual_seed(s)  # noqa: E731
    img = torch.rand(T, 3, 256, 256, generator=g(91)) * 2 - 1
    flow = torch.rand(T, 1, 1, generator=g(92)) * 3          # [B, len_seq = 1, f_dim]   (dataset_wlbl.py:567-570)
    weather = torch.randn(T, 1, 16, generator=g(93))          # [B, len_seq = 1, w_dim]
    time = torch.rand(T, 1, generator=g(94)) * 2 + torch.arange(T).float()[:, None] * 0.01
    return img, flow, weather, time


def main():
    import gen_golden as GG
    GG.install_shims()
    from oracle import autoencoder as OA
    from oracle import conditioner as OC
    from oracle import unet as OU
    from oracle import vq as OV
    sys.modules["taming.modules.vqvae.quantize"].VectorQuantizer2 = OV.VectorQuantizer2
    torch.cuda.current_device = lambda: "cpu"
    import torchvision
    _resnet50 = torchvision.models.resnet50
    torchvision.models.resnet50 = lambda pretrained=False, **k: _resnet50(weights=None)
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    from ldm.models.autoencoder import VQModelInterface
    from STDiff.models import UnetCond

    fs = GG.load_cfg("configs/latent-diffusion/stdiff_cin-ldm-vq-f8.yaml")["model"]["params"]["first_stage_config"]["params"]
    dd, n_embed, embed_dim = fs["ddconfig"], fs["n_embed"], fs["embed_dim"]
    first = VQModelInterface(embed_dim=embed_dim, n_embed=n_embed, ddconfig=dd, lossconfig={"target": "torch.nn.Identity"})
    first.load_state_dict(OU.synthetic_state_dict(OA.vq_param_shapes(dd, embed_dim, n_embed), seed=4), strict=True)
    first.eval()

    m = UnetCond(device="cpu", cond_args=_Args(OC.COND_ARGS))
    m.convs = first                                           # ddpm.py:535-536
    ref_names = [(k, tuple(v.shape)) for k, v in m.state_dict().items() if not k.startswith("convs.")]
    assert ref_names == OC.param_shapes(), "conditioner: oracle parameter inventory differs from the reference"
    sd = OC.synthetic_state_dict()
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith("convs.") for k in missing), (missing, unexpected)

    img, flow, weather, time = make_inputs()
    mixed = (img, flow, weather, time)                        # squeeze(0) in UnetCond.forward is a no-op for B > 1
    out = {"T": T, "flow": flow, "weather": weather, "time": time, "img_seed": 91}
    with torch.no_grad():
        out["z"] = first.encoder(img)
        m.eval()
        out["context_eval"] = m(mixed)
        m.conv_cat[1].train()                                 # BatchNorm on batch statistics, Dropout still off
        out["context_bn_train"] = m(mixed)
        m.eval()
        # intermediates, from the reference's own sub-modules
        c = next(iter(m.cond_xform(time, broadcast=True).unbind(dim=1)))
        out["fourier"] = c
        out["time_style"] = m.scaled_styles(c=c)
        out["flow_style"] = m.f_mlp(flow, "train")
        out["weather_style"] = m.w_mlp(weather, "train")
        seq = torch.randn(3, 4, 16, generator=torch.Generator().manual_seed(95))   # a longer sequence: the recurrence
        out["lstm_seq_in"], out["lstm_seq_out"] = seq, m.w_mlp(seq, "train")
    path = os.path.join(ROOT, "tests", "golden", "conditioner.pt")
    torch.save(out, path)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB); context {tuple(out['context_eval'].shape)}, "
          f"|context| = {float(out['context_eval'].norm()):.3f}")
    # the restatement against the reference, right here
    o = OC.unet_cond_forward(sd, out["z"], flow, weather, time, bn_training=False)
    ob = OC.unet_cond_forward(sd, out["z"], flow, weather, time, bn_training=True)
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())  # noqa: E731
    print("oracle vs reference: eval %.2e, bn-train %.2e, styles %.2e %.2e %.2e, fourier %.2e" % (
        rel(o["context"], out["context_eval"]), rel(ob["context"], out["context_bn_train"]),
        rel(o["time_style"], out["time_style"]), rel(o["flow_style"], out["flow_style"]),
        rel(o["weather_style"], out["weather_style"]), rel(OC.fourier_features(time), out["fourier"])))
    print("oracle LSTM recurrence (S = 4): %.2e" % rel(OC.lstm_mlp(sd, "w_mlp", out["lstm_seq_in"]), out["lstm_seq_out"]))


if __name__ == "__main__":
    main()
