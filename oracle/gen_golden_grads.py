"""TEST INFRASTRUCTURE ONLY -- golden GRADIENTS of the training step, made by the unmodified reference.

Runs the reference `LatentDiffusion.p_losses` (ldm/models/diffusion/ddpm.py:1036-1078) followed by
`loss.backward()` through the reference UNet (openaimodel.py / attention.py, including its
`checkpoint` recomputation) on CPU fp32, on the same seeded inputs and synthetic weights as the
`p_losses.pt` golden, and stores for every one of the 626 parameters
    norm[name] = ||grad||_2            proj[name] = <grad, r_i>,  r_i = randn(shape, seed 1000 + i)
(two numbers per tensor: the full gradients are 1.6 GB), plus a few small gradients in full and the
gradient with respect to the conditioning.  tests/test_train_gpu.py regenerates r_i from the seeds.

Run in the build container (needs /root/reference):  python oracle/gen_golden_grads.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

FULL = ["time_embed.0.bias", "time_embed.2.bias", "out.2.weight", "out.0.weight", "input_blocks.0.0.weight",
        "input_blocks.1.0.emb_layers.1.bias", "input_blocks.1.1.transformer_blocks.0.norm2.weight",
        "input_blocks.4.0.skip_connection.weight", "middle_block.1.transformer_blocks.0.attn2.to_k.weight",
        "output_blocks.8.1.proj_out.bias", "input_blocks.3.0.op.bias"]


def direction(i, shape):
    return torch.randn(shape, generator=torch.Generator().manual_seed(1000 + i))


def main():
    import gen_golden as GG
    GG.install_shims()
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    from oracle import unet as OU
    from ldm.models.diffusion.ddpm import LatentDiffusion

    cfg_st = GG.load_cfg("configs/latent-diffusion/stdiff_cin-ldm-vq-f8.yaml")["model"]["params"]
    p = dict(cfg_st)
    p["first_stage_config"] = {"target": "ldm.models.autoencoder.IdentityFirstStage"}
    p["cond_stage_config"] = {"target": "torch.nn.Identity"}
    p["cond_stage_trainable"] = False
    p.pop("cond_stage_key", None)
    p["use_ema"] = False
    ld = LatentDiffusion(**p)
    ucfg = cfg_st["unet_config"]["params"]
    sd = OU.synthetic_state_dict(OU.unet_param_shapes(ucfg), seed=2)
    unet = ld.model.diffusion_model
    unet.load_state_dict(sd, strict=True)
    ld.train()

    g = GG.g
    x0 = torch.randn(2, 4, 32, 32, generator=g(51))
    noise = torch.randn(2, 4, 32, 32, generator=g(52))
    t = torch.tensor([10, 700])
    c2 = torch.randn(4, 4, 512, generator=g(53)).requires_grad_(True)
    loss, _ = ld.p_losses(x0, c2, t, noise=noise)
    loss.backward()
    out = {"loss": loss.detach(), "norm": {}, "proj": {}, "full": {}, "dcond": c2.grad.clone(),
           "names": [n for n, _ in unet.named_parameters()]}
    for i, (name, prm) in enumerate(unet.named_parameters()):
        gr = prm.grad
        assert gr is not None, name
        out["norm"][name] = float(gr.double().norm())
        out["proj"][name] = float((gr.double() * direction(i, gr.shape).double()).sum())
        if name in FULL:
            out["full"][name] = gr.clone()
    path = os.path.join(ROOT, "tests", "golden", "p_losses_grads.pt")
    torch.save(out, path)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB); loss = {float(loss):.6f}")


if __name__ == "__main__":
    main()
