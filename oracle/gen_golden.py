"""Generate tests/golden/*.pt by running the UNMODIFIED reference modules from /root/reference.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Runs only in the build container (the GPU box has no
/root/reference); the outputs are committed.  Usage:  python -m oracle.gen_golden [--out tests/golden]

Import shims: pytorch_lightning (LightningModule = nn.Module + .device),
omegaconf.listconfig.ListConfig, taming VectorQuantizer2 stub, torchvision.utils.make_grid stub.
Weights: oracle.unet.synthetic_state_dict (reference-independent, deterministic) loaded with
load_state_dict(strict=True) into the reference's own constructors, which also proves that the
parameter names/shapes of oracle.*_param_shapes equal the reference's.
"""
from __future__ import annotations

import argparse
import os
import sys
import types

import numpy as np
import torch
import yaml

REF = os.environ.get("EALDM_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def install_shims():
    import torch.nn as nn

    pl = types.ModuleType("pytorch_lightning")

    class LightningModule(nn.Module):
        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")

        def log(self, *a, **k):
            pass

        def log_dict(self, *a, **k):
            pass

    pl.LightningModule = LightningModule
    util = types.ModuleType("pytorch_lightning.utilities")
    dist = types.ModuleType("pytorch_lightning.utilities.distributed")
    dist.rank_zero_only = lambda f: f
    util.distributed = dist
    pl.utilities = util
    sys.modules.update({"pytorch_lightning": pl, "pytorch_lightning.utilities": util,
                        "pytorch_lightning.utilities.distributed": dist})

    oc = types.ModuleType("omegaconf")
    lc = types.ModuleType("omegaconf.listconfig")
    lc.ListConfig = type("ListConfig", (list,), {})
    oc.listconfig = lc
    sys.modules.update({"omegaconf": oc, "omegaconf.listconfig": lc})

    names = ["taming", "taming.modules", "taming.modules.vqvae", "taming.modules.vqvae.quantize"]
    mods = {n: types.ModuleType(n) for n in names}
    mods[names[-1]].VectorQuantizer2 = type("VectorQuantizer2", (nn.Module,), {})
    sys.modules.update(mods)

    try:
        import torchvision.utils  # noqa: F401
    except Exception:
        tv = types.ModuleType("torchvision")
        tvu = types.ModuleType("torchvision.utils")
        tvu.make_grid = lambda *a, **k: None
        tv.utils = tvu
        sys.modules.update({"torchvision": tv, "torchvision.utils": tvu})
    if REF not in sys.path:
        sys.path.insert(0, REF)


def load_cfg(rel):
    with open(os.path.join(REF, rel)) as f:
        return yaml.safe_load(f)


def g(seed):
    return torch.Generator().manual_seed(seed)


def golden_module_inputs():
    """Seeded inputs of the module-level golden vectors (shared with tests/test_oracle_golden.py)."""
    return {
        "res_in4": {"x": torch.randn(1, 256, 16, 16, generator=g(14)), "emb": torch.randn(1, 1024, generator=g(13))},
        "st_in4": {"x": torch.randn(1, 512, 16, 16, generator=g(15)), "context": torch.randn(1, 4, 512, generator=g(16))},
        "down_in3": {"x": torch.randn(1, 256, 32, 32, generator=g(17))},
        "up_out2": {"x": torch.randn(1, 1024, 8, 8, generator=g(18))},
        "attnblock_in4": {"x": torch.randn(1, 512, 16, 16, generator=g(19))},
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    sys.path.insert(0, ROOT)
    install_shims()
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())

    from oracle import autoencoder as OA
    from oracle import unet as OU

    from ldm.models.diffusion.ddim import DDIMSampler
    from ldm.models.diffusion.ddpm import LatentDiffusion
    from ldm.modules.diffusionmodules.openaimodel import UNetModel
    from ldm.modules.diffusionmodules.util import timestep_embedding
    from ldm.models.autoencoder import AutoencoderKL

    def save(name, obj):
        path = os.path.join(args.out, name)
        torch.save(obj, path)
        print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")

    # ---- (i) schedules + (ii) timestep embedding ------------------------------------------------
    cfg_st = load_cfg("configs/latent-diffusion/stdiff_cin-ldm-vq-f8.yaml")["model"]["params"]
    cfg_un = load_cfg("configs/latent-diffusion/uncond_cin-ldm-vq-f8.yaml")["model"]["params"]

    def make_ld(params, with_ctx):
        p = dict(params)
        p["first_stage_config"] = {"target": "ldm.models.autoencoder.IdentityFirstStage"}
        p["cond_stage_config"] = {"target": "torch.nn.Identity"} if with_ctx else "__is_unconditional__"
        p["cond_stage_trainable"] = False
        p.pop("cond_stage_key", None)
        p["use_ema"] = False
        return LatentDiffusion(**p)

    ld_un = make_ld(cfg_un, False)
    ld_st = make_ld(cfg_st, True)
    sched = {"register": {k: v.clone() for k, v in ld_st.named_buffers() if "." not in k}}
    sched["register"]["lvlb_weights"] = ld_st.lvlb_weights.clone()
    for S in (10, 50):
        for eta in (0.0, 1.0):
            s = DDIMSampler(ld_st)
            s.register_buffer = lambda n, a, _s=s: setattr(_s, n, a)  # ddim.py:18-22 hard-codes cuda
            s.make_schedule(S, ddim_eta=eta, verbose=False)
            sched[f"S{S}_eta{eta}"] = {
                "ddim_timesteps": torch.as_tensor(np.asarray(s.ddim_timesteps)),
                "ddim_alphas": torch.as_tensor(np.asarray(s.ddim_alphas), dtype=torch.float64),
                "ddim_alphas_prev": torch.as_tensor(np.asarray(s.ddim_alphas_prev), dtype=torch.float64),
                "ddim_sigmas": torch.as_tensor(np.asarray(s.ddim_sigmas), dtype=torch.float64),
                "ddim_sqrt_one_minus_alphas": torch.as_tensor(np.asarray(s.ddim_sqrt_one_minus_alphas),
                                                              dtype=torch.float64),
                "dtypes": {k: str(getattr(getattr(s, k), "dtype", type(getattr(s, k))))
                           for k in ("ddim_alphas", "ddim_alphas_prev", "ddim_sigmas",
                                     "ddim_sqrt_one_minus_alphas")},
            }
    sched["timestep_embedding"] = {
        "t": torch.tensor([1, 501, 981, 0, 999]),
        "emb": timestep_embedding(torch.tensor([1, 501, 981, 0, 999]), 256),
    }
    save("schedule.pt", sched)

    # ---- (iv) full UNet forward, both configs, B=2, synthetic weights ----------------------------
    ucfg_un = cfg_un["unet_config"]["params"]
    ucfg_st = cfg_st["unet_config"]["params"]
    unets = {}
    for name, ucfg, ld in (("uncond", ucfg_un, ld_un), ("stdiff", ucfg_st, ld_st)):
        shapes = OU.unet_param_shapes(ucfg)
        ref_names = [(k, tuple(v.shape)) for k, v in ld.model.diffusion_model.state_dict().items()]
        assert ref_names == shapes, f"{name}: oracle parameter inventory differs from the reference"
        sd = OU.synthetic_state_dict(shapes, seed=1 if name == "uncond" else 2)
        ld.model.diffusion_model.load_state_dict(sd, strict=True)
        ld.eval()
        unets[name] = (ucfg, sd, ld)
        x = torch.randn(2, 4, 32, 32, generator=g(11))
        t = torch.tensor([981, 21])
        ctx = torch.randn(2, 4, 512, generator=g(12)) if name == "stdiff" else None
        with torch.no_grad():
            if ctx is None:
                eps = ld.apply_model(x, t, None) if False else ld.model.diffusion_model(x, t)
            else:
                eps = ld.apply_model(x, t, ctx)
        save(f"unet_{name}_fwd.pt", {"x": x, "t": t, "context": ctx, "eps": eps,
                                     "n_params": sum(v.numel() for v in sd.values())})

    # ---- (iii) a few module-level vectors out of the stdiff / uncond nets ------------------------
    # inputs are regenerated by the tests from the same seeds (golden_module_inputs); only the
    # reference outputs are stored.
    mods = {}
    um = unets["stdiff"][2].model.diffusion_model
    uu = unets["uncond"][2].model.diffusion_model
    mi = golden_module_inputs()
    with torch.no_grad():
        mods["res_in4"] = um.input_blocks[4][0](mi["res_in4"]["x"], mi["res_in4"]["emb"])  # 256->512 + skip conv
        mods["st_in4"] = um.input_blocks[4][1](mi["st_in4"]["x"], mi["st_in4"]["context"])
        mods["down_in3"] = um.input_blocks[3][0](mi["down_in3"]["x"])
        mods["up_out2"] = um.output_blocks[2][2](mi["up_out2"]["x"])
        mods["attnblock_in4"] = uu.input_blocks[4][1](mi["attnblock_in4"]["x"])
    save("unet_modules.pt", mods)

    # ---- (v) 10-step DDIM trajectories -----------------------------------------------------------
    def run_ddim(ld, B, S, eta, cond, uc, ugs, seed):
        sampler = DDIMSampler(ld)
        sampler.register_buffer = lambda n, a, _s=sampler: setattr(_s, n, a)
        x_T = torch.randn(B, 4, 32, 32, generator=g(seed))
        trace = {"e_t": [], "x_prev": [], "pred_x0": [], "noise": []}
        orig_apply = ld.apply_model
        import ldm.models.diffusion.ddim as ddim_mod
        orig_noise_like = ddim_mod.noise_like
        gen_noise = g(seed + 1)

        def noise_like(shape, device, repeat=False):
            n = torch.randn(shape, generator=gen_noise)
            trace["noise"].append(n)
            return n

        ddim_mod.noise_like = noise_like
        orig_p = sampler.p_sample_ddim

        def p_wrap(*a, **k):
            out = orig_p(*a, **k)
            trace["x_prev"].append(out[0])
            trace["pred_x0"].append(out[1])
            return out

        sampler.p_sample_ddim = p_wrap

        def apply_wrap(x, t, c, **k):
            e = orig_apply(x, t, c, **k)
            trace["raw_eps_last"] = e
            return e

        ld.apply_model = apply_wrap
        try:
            with torch.no_grad():
                samples, _ = sampler.sample(S=S, batch_size=B, shape=(4, 32, 32), conditioning=cond, eta=eta,
                                            x_T=x_T, verbose=False, unconditional_guidance_scale=ugs,
                                            unconditional_conditioning=uc)
        finally:
            ddim_mod.noise_like = orig_noise_like
            ld.apply_model = orig_apply
        # recover the guided eps of every step from pred_x0: not needed -- x_prev/pred_x0 pin it
        out = {"x_T": x_T, "samples": samples, "x_prev": torch.stack(trace["x_prev"]),
               "pred_x0": torch.stack(trace["pred_x0"])}
        if eta != 0.0:
            out["noise"] = torch.stack(trace["noise"])
        else:
            out["noise_seed"] = seed + 1  # drawn (ddim.py:200) but multiplied by sigma = 0
        return out

    ld = unets["uncond"][2]
    traj = {"config1_uncond_B4_S10_eta0": dict(run_ddim(ld, 4, 10, 0.0, None, None, 1.0, 21), S=10, eta=0.0, ugs=1.0)}
    ld = unets["stdiff"][2]
    cond = torch.randn(2, 4, 512, generator=g(31))
    uc = torch.randn(2, 4, 512, generator=g(32))
    traj["stdiff_B2_S10_eta1_cfg2"] = dict(run_ddim(ld, 2, 10, 1.0, cond, uc, 2.0, 41), S=10, eta=1.0, ugs=2.0,
                                           cond=cond, uc=uc)
    save("ddim_traj.pt", traj)

    # ---- (vi) p_losses ---------------------------------------------------------------------------
    ld = unets["stdiff"][2]
    x0 = torch.randn(2, 4, 32, 32, generator=g(51))
    noise = torch.randn(2, 4, 32, 32, generator=g(52))
    t = torch.tensor([10, 700])
    c2 = torch.randn(4, 4, 512, generator=g(53))  # [c_neg; c], ddpm.py:891-893
    with torch.no_grad():
        loss, ld_dict = ld.p_losses(x0, c2, t, noise=noise)
        xq = ld.q_sample(x0, t, noise)
    save("p_losses.pt", {"x0": x0, "noise": noise, "t": t, "cond2": c2, "loss": loss, "q_sample": xq,
                         "loss_dict": {k: v.clone() for k, v in ld_dict.items()}})

    # ---- (vii) AutoencoderKL encode / decode at 128x128, B=1 -------------------------------------
    acfg = load_cfg("configs/autoencoder/autoencoder_kl_32x32x4.yaml")["model"]["params"]
    ae = AutoencoderKL(ddconfig=acfg["ddconfig"], lossconfig={"target": "torch.nn.Identity"},
                       embed_dim=acfg["embed_dim"])
    shapes = OA.autoencoder_kl_param_shapes(acfg["ddconfig"], acfg["embed_dim"])
    ref_names = [(k, tuple(v.shape)) for k, v in ae.state_dict().items()]
    assert ref_names == shapes, "autoencoder: oracle parameter inventory differs from the reference"
    sd = OU.synthetic_state_dict(shapes, seed=3)
    ae.load_state_dict(sd, strict=True)
    ae.eval()
    img = torch.rand(1, 3, 128, 128, generator=g(61)) * 2 - 1
    z = torch.randn(1, 4, 16, 16, generator=g(62))
    with torch.no_grad():
        post = ae.encode(img)
        dec = ae.decode(z)
    save("autoencoder_kl.pt", {"img": img, "moments": post.parameters, "mean": post.mean, "std": post.std,
                               "z": z, "dec": dec})
    print("done")


if __name__ == "__main__":
    main()
