"""TEST INFRASTRUCTURE ONLY -- golden DDPM ancestral sampling steps made by the unmodified reference
`LatentDiffusion.p_sample_loop` (ldm/models/diffusion/ddpm.py:1190-1247) on CPU fp32: stdiff UNet, B=2, the last 6
timesteps (5..0, so the t == 0 no-noise branch is covered), noise captured from the reference's own `noise_like` calls.
-> tests/golden/ddpm_ancestral.pt.  Run in the build container:  python oracle/gen_golden_ddpm.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    import gen_golden as GG
    GG.install_shims()
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    from oracle import unet as OU
    import ldm.models.diffusion.ddpm as ddpm_mod
    from ldm.models.diffusion.ddpm import LatentDiffusion

    params = dict(GG.load_cfg("configs/latent-diffusion/stdiff_cin-ldm-vq-f8.yaml")["model"]["params"])
    params["first_stage_config"] = {"target": "ldm.models.autoencoder.IdentityFirstStage"}
    params["cond_stage_config"] = {"target": "torch.nn.Identity"}
    params["cond_stage_trainable"] = False
    params.pop("cond_stage_key", None)
    params["use_ema"] = False
    ld = LatentDiffusion(**params)
    ucfg = params["unet_config"]["params"]
    ld.model.diffusion_model.load_state_dict(OU.synthetic_state_dict(OU.unet_param_shapes(ucfg), seed=2), strict=True)
    ld.eval()
    x_T = torch.randn(2, 4, 32, 32, generator=GG.g(91)) * 0.8
    cond = torch.randn(2, 4, 512, generator=GG.g(92))
    gen = GG.g(93)
    noises, imgs = [], []
    orig = ddpm_mod.noise_like

    def noise_like(shape, device, repeat=False):
        n = torch.randn(shape, generator=gen)
        noises.append(n)
        return n

    ddpm_mod.noise_like = noise_like
    try:
        with torch.no_grad():
            out = ld.p_sample_loop(cond, (2, 4, 32, 32), x_T=x_T, verbose=False, timesteps=6,
                                   img_callback=lambda img, i: imgs.append(img.clone()))
    finally:
        ddpm_mod.noise_like = orig
    path = os.path.join(ROOT, "tests", "golden", "ddpm_ancestral.pt")
    torch.save({"x_T": x_T, "cond": cond, "noise": torch.stack(noises), "imgs": torch.stack(imgs), "out": out,
                "clip_denoised": bool(ld.clip_denoised)}, path)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB); clip_denoised = {ld.clip_denoised}")


if __name__ == "__main__":
    main()
