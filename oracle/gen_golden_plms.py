"""TEST INFRASTRUCTURE ONLY -- golden PLMS trajectories made by the unmodified reference `PLMSSampler`
(ldm/models/diffusion/plms.py) on CPU fp32, with the same synthetic weights as the other golden files:
  uncond_cin UNet, B=2, S=10;  stdiff UNet with classifier-free guidance 2.0, B=2, S=8.
Stores x_T, conditioning, and per step x_prev / pred_x0 (-> tests/golden/plms_traj.pt).
Run in the build container (needs /root/reference):  python oracle/gen_golden_plms.py
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
    from ldm.models.diffusion.ddpm import LatentDiffusion
    from ldm.models.diffusion.plms import PLMSSampler

    def make_ld(rel, with_ctx, seed):
        params = dict(GG.load_cfg(rel)["model"]["params"])
        params["first_stage_config"] = {"target": "ldm.models.autoencoder.IdentityFirstStage"}
        params["cond_stage_config"] = {"target": "torch.nn.Identity"} if with_ctx else "__is_unconditional__"
        params["cond_stage_trainable"] = False
        params.pop("cond_stage_key", None)
        params["use_ema"] = False
        ld = LatentDiffusion(**params)
        ucfg = params["unet_config"]["params"]
        ld.model.diffusion_model.load_state_dict(OU.synthetic_state_dict(OU.unet_param_shapes(ucfg), seed=seed), strict=True)
        return ld.eval()

    def run(ld, B, S, cond, uc, ugs, seed):
        sampler = PLMSSampler(ld)
        sampler.register_buffer = lambda n, a, _s=sampler: setattr(_s, n, a)   # plms.py:18-22 hard-codes cuda
        x_T = torch.randn(B, 4, 32, 32, generator=GG.g(seed))
        xs, ps = [], []
        orig = sampler.p_sample_plms

        def wrap(*a, **k):
            out = orig(*a, **k)
            xs.append(out[0]); ps.append(out[1])
            return out

        sampler.p_sample_plms = wrap
        with torch.no_grad():
            samples, _ = sampler.sample(S=S, batch_size=B, shape=(4, 32, 32), conditioning=cond, eta=0.0, x_T=x_T,
                                        verbose=False, unconditional_guidance_scale=ugs, unconditional_conditioning=uc)
        return {"x_T": x_T, "samples": samples, "x_prev": torch.stack(xs), "pred_x0": torch.stack(ps), "S": S, "ugs": ugs,
                "cond": cond, "uc": uc}

    out = {}
    out["uncond_B2_S10"] = run(make_ld("configs/latent-diffusion/uncond_cin-ldm-vq-f8.yaml", False, 1), 2, 10, None, None, 1.0, 71)
    cond = torch.randn(2, 4, 512, generator=GG.g(72))
    uc = torch.randn(2, 4, 512, generator=GG.g(73))
    out["stdiff_B2_S8_cfg2"] = run(make_ld("configs/latent-diffusion/stdiff_cin-ldm-vq-f8.yaml", True, 2), 2, 8, cond, uc, 2.0, 74)
    path = os.path.join(ROOT, "tests", "golden", "plms_traj.pt")
    torch.save(out, path)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
