"""`LatentDiffusion` / `DiffusionWrapper` with the reference's public contract for the hot path
(reference: ldm/models/diffusion/ddpm.py -- DDPM :46-171,276-294; LatentDiffusion :428-566,713-771,
878-921,1011-1078,1267-1276; DiffusionWrapper :1443-1469).

This is a plain `nn.Module` (the reference derives from pytorch_lightning.LightningModule, which is
orchestration only): it owns the schedule buffers, routes conditioning to the UNet, and implements
`q_sample`, `p_losses`, `apply_model`, `encode/decode_first_stage` and `sample_log` with the same
arithmetic, including EALDM's deviations: classifier-free guidance inside the training loss with a
hard-coded scale of 2 (ddpm.py:442,1040-1044).  Trainer hooks, logging and dataset plumbing stay in
the reference (out of scope).
"""
from __future__ import annotations

import copy
from contextlib import contextmanager
from functools import partial

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .ddim import DDIMSampler
from .ema import LitEma, invalidate_packed_modules
from .util import count_params, instantiate_from_config, make_beta_schedule, revalidate_packed


def disabled_train(self, mode=True):
    return self


class IdentityFirstStage(nn.Module):
    """ldm/models/autoencoder.py:426-443"""

    def __init__(self, *args, vq_interface=False, **kwargs):
        super().__init__()
        self.vq_interface = vq_interface

    def encode(self, x, *args, **kwargs):
        return x

    def decode(self, x, *args, **kwargs):
        return x

    def quantize(self, x, *args, **kwargs):
        return (x, None, [None, None, None]) if self.vq_interface else x

    def forward(self, x, *args, **kwargs):
        return x


class _CfgMse(torch.autograd.Function):
    """loss_simple[b] = mean((e_u + s (e_c - e_u) - target)^2) as one kernel each way (ddpm.py:1040-1060)."""

    @staticmethod
    def forward(ctx, e_u, e_c, target, s):
        e_c, target = e_c.contiguous(), target.contiguous()
        e_u = None if e_u is None else e_u.contiguous()
        ctx.save_for_backward(e_c, target, *(() if e_u is None else (e_u,)))
        ctx.s = s
        return ops.cfg_mse(e_c, target, e_uncond=e_u, cfg_scale=s)

    @staticmethod
    def backward(ctx, w):
        e_c, target, *rest = ctx.saved_tensors
        e_u = rest[0] if rest else None
        de_u, de_c = ops.cfg_mse_bwd(e_c, target, w.float().contiguous(), e_uncond=e_u, cfg_scale=ctx.s)
        return de_u, de_c, None, None


class DiffusionWrapper(nn.Module):
    """ddpm.py:1443-1469"""

    def __init__(self, diff_model_config, conditioning_key):
        super().__init__()
        self.diffusion_model = instantiate_from_config(diff_model_config)
        self.conditioning_key = conditioning_key
        assert self.conditioning_key in [None, "concat", "crossattn", "hybrid", "adm"]

    def forward(self, x, t, c_concat: list = None, c_crossattn: list = None):
        key = self.conditioning_key
        if key is None:
            return self.diffusion_model(x, t)
        if key == "concat":
            return self.diffusion_model(torch.cat([x] + c_concat, dim=1), t)
        if key == "crossattn":
            cc = c_crossattn[0] if len(c_crossattn) == 1 else torch.cat(c_crossattn, 1)
            return self.diffusion_model(x, t, context=cc)
        if key == "hybrid":
            return self.diffusion_model(torch.cat([x] + c_concat, dim=1), t, context=torch.cat(c_crossattn, 1))
        if key == "adm":
            return self.diffusion_model(x, t, y=c_crossattn[0])
        raise NotImplementedError()


class LatentDiffusion(nn.Module):
    def __init__(self, unet_config, first_stage_config=None, cond_stage_config="__is_unconditional__",
                 num_timesteps_cond=None, cond_stage_key="image", cond_stage_trainable=False, concat_mode=True,
                 cond_stage_forward=None, conditioning_key=None, scale_factor=1.0, scale_by_std=False,
                 timesteps=1000, beta_schedule="linear", loss_type="l2", ckpt_path=None, ignore_keys=(),
                 monitor="val/loss", use_ema=False, first_stage_key="image", image_size=256, channels=3,
                 log_every_t=100, clip_denoised=True, linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3,
                 given_betas=None, original_elbo_weight=0., v_posterior=0., l_simple_weight=1.,
                 parameterization="eps", scheduler_config=None, use_positional_encodings=False,
                 learn_logvar=False, logvar_init=0., base_learning_rate=None, **unused):
        super().__init__()
        assert parameterization in ["eps", "x0"], 'currently only supporting "eps" and "x0"'
        self.unconditional_guidance_scale = 2.  # ddpm.py:442 (hard-coded in the reference)
        self.parameterization = parameterization
        self.num_timesteps_cond = 1 if num_timesteps_cond is None else num_timesteps_cond
        self.scale_by_std = scale_by_std
        if conditioning_key is None:
            conditioning_key = "concat" if concat_mode else "crossattn"
        if cond_stage_config == "__is_unconditional__":
            conditioning_key = None
        self.cond_stage_model = None
        self.clip_denoised = False
        self.log_every_t = log_every_t
        self.first_stage_key, self.cond_stage_key = first_stage_key, cond_stage_key
        self.image_size, self.channels = image_size, channels
        self.model = DiffusionWrapper(unet_config, conditioning_key)
        count_params(self.model, verbose=False)
        self.use_ema = use_ema          # NB: the reference's default is True (ddpm.py:57); its shipped yaml leaves it on
        if use_ema:                     # `model_ema.*` buffers in the reference's checkpoint layout (ema.py:5-23)
            self.model_ema = LitEma(self.model)
        self.use_scheduler = scheduler_config is not None
        if self.use_scheduler:
            self.scheduler_config = scheduler_config
        self.learning_rate = base_learning_rate      # main.py:741-745 overwrites it with ngpu * bs * base_lr
        self.shorten_cond_schedule = self.num_timesteps_cond > 1
        if self.shorten_cond_schedule:
            raise NotImplementedError("num_timesteps_cond > 1 (the shortened conditioning schedule, ddpm.py:497-505)")
        self.v_posterior, self.original_elbo_weight, self.l_simple_weight = v_posterior, original_elbo_weight, l_simple_weight
        self.monitor = monitor
        self.loss_type = loss_type
        self.register_schedule(given_betas=given_betas, beta_schedule=beta_schedule, timesteps=timesteps,
                               linear_start=linear_start, linear_end=linear_end, cosine_s=cosine_s)
        self.learn_logvar = learn_logvar
        self.logvar = torch.full(fill_value=logvar_init, size=(self.num_timesteps,))
        if learn_logvar:
            self.logvar = nn.Parameter(self.logvar, requires_grad=True)
        self.concat_mode, self.cond_stage_trainable = concat_mode, cond_stage_trainable
        self.scale_factor = scale_factor
        self.cond_stage_forward = cond_stage_forward
        self.instantiate_first_stage(first_stage_config)
        self.instantiate_cond_stage(cond_stage_config)
        if ckpt_path is not None:
            self.init_from_ckpt(ckpt_path, list(ignore_keys))

    # ---- construction -----------------------------------------------------------------------------
    @property
    def device(self):
        return self.betas.device

    def register_schedule(self, given_betas=None, beta_schedule="linear", timesteps=1000, linear_start=1e-4,
                          linear_end=2e-2, cosine_s=8e-3):
        """ddpm.py:119-171: float64 numpy math, float32 buffers."""
        betas = given_betas if given_betas is not None else make_beta_schedule(
            beta_schedule, timesteps, linear_start=linear_start, linear_end=linear_end, cosine_s=cosine_s)
        alphas = 1. - betas
        ac = np.cumprod(alphas, axis=0)
        ac_prev = np.append(1., ac[:-1])
        self.num_timesteps = int(betas.shape[0])
        self.linear_start, self.linear_end = linear_start, linear_end
        f32 = partial(torch.tensor, dtype=torch.float32)
        reg = self.register_buffer
        reg("betas", f32(betas))
        reg("alphas_cumprod", f32(ac))
        reg("alphas_cumprod_prev", f32(ac_prev))
        reg("sqrt_alphas_cumprod", f32(np.sqrt(ac)))
        reg("sqrt_one_minus_alphas_cumprod", f32(np.sqrt(1. - ac)))
        reg("log_one_minus_alphas_cumprod", f32(np.log(1. - ac)))
        reg("sqrt_recip_alphas_cumprod", f32(np.sqrt(1. / ac)))
        reg("sqrt_recipm1_alphas_cumprod", f32(np.sqrt(1. / ac - 1)))
        post_var = (1 - self.v_posterior) * betas * (1. - ac_prev) / (1. - ac) + self.v_posterior * betas
        reg("posterior_variance", f32(post_var))
        reg("posterior_log_variance_clipped", f32(np.log(np.maximum(post_var, 1e-20))))
        reg("posterior_mean_coef1", f32(betas * np.sqrt(ac_prev) / (1. - ac)))
        reg("posterior_mean_coef2", f32((1. - ac_prev) * np.sqrt(alphas) / (1. - ac)))
        if self.parameterization == "eps":
            lvlb = self.betas ** 2 / (2 * self.posterior_variance * f32(alphas) * (1 - self.alphas_cumprod))
        else:
            lvlb = 0.5 * np.sqrt(torch.Tensor(ac)) / (2. * 1 - torch.Tensor(ac))
        lvlb[0] = lvlb[1]
        reg("lvlb_weights", lvlb, persistent=False)
        assert not torch.isnan(self.lvlb_weights).all()

    def instantiate_first_stage(self, config):
        model = IdentityFirstStage() if config is None else instantiate_from_config(config)
        self.first_stage_model = model.eval()
        self.first_stage_model.train = disabled_train.__get__(self.first_stage_model)
        for p in self.first_stage_model.parameters():
            p.requires_grad = False

    def instantiate_cond_stage(self, config):
        if config in ("__is_unconditional__", None):
            self.cond_stage_model = None
        elif config == "__is_first_stage__":
            self.cond_stage_model = self.first_stage_model
        else:
            self.cond_stage_model = instantiate_from_config(config)
            if not self.cond_stage_trainable:
                self.cond_stage_model.eval()
                for p in self.cond_stage_model.parameters():
                    p.requires_grad = False
            if self.cond_stage_key == "mixed" and hasattr(self.cond_stage_model, "convs"):
                # the EALDM conditioner encodes its frames with the first stage (ddpm.py:535-536)
                self.cond_stage_model.convs = self.first_stage_model

    def init_from_ckpt(self, path, ignore_keys=(), only_model=False):
        """ddpm.py:188-201.  With `use_ema=True` the checkpoint's `model_ema.*` buffers load into `self.model_ema`; with
        `use_ema=False` they would be dropped by strict=False, so say so instead of sampling raw weights silently."""
        sd = torch.load(path, map_location="cpu")
        sd = sd.get("state_dict", sd)
        for k in list(sd.keys()):
            if any(k.startswith(ik) for ik in ignore_keys):
                del sd[k]
        target = self.model if only_model else self
        missing, unexpected = target.load_state_dict(sd, strict=False)
        print(f"Restored from {path} with {len(missing)} missing and {len(unexpected)} unexpected keys")
        if not self.use_ema and any(k.startswith("model_ema.") for k in unexpected):
            print("WARNING: the checkpoint holds EMA weights (model_ema.*) but use_ema=False: they were NOT loaded and "
                  "ema_scope() is a no-op; construct the model with use_ema=True to sample under the EMA weights")

    @contextmanager
    def ema_scope(self, context=None):
        """ddpm.py:173-186: run the body under the EMA weights.  The packed kernel-layout weights (and CUDA graphs) of
        the UNet are dropped on both switches (ema.LitEma.copy_to / restore)."""
        if self.use_ema:
            self.model_ema.store(self.model.parameters())
            self.model_ema.copy_to(self.model)
            if context is not None:
                print(f"{context}: Switched to EMA weights")
        try:
            yield None
        finally:
            if self.use_ema:
                self.model_ema.restore(self.model.parameters(), self.model)
                if context is not None:
                    print(f"{context}: Restored training weights")

    def on_train_batch_end(self, *args, **kwargs):
        """ddpm.py:370-372 (a no-op once the shadows are bound to optim.FusedAdamWEMA, whose kernel updates them)."""
        if self.use_ema and self.model_ema._fused is None:
            self.model_ema(self.model)

    def configure_optimizers(self):
        """ddpm.py:1409-1431: AdamW over the UNet (+ the conditioner when trainable, + logvar when learned)."""
        params = list(self.model.parameters())
        if self.cond_stage_trainable:
            params = params + list(self.cond_stage_model.parameters())
        if self.learn_logvar:
            params.append(self.logvar)
        opt = torch.optim.AdamW(params, lr=self.learning_rate)
        if self.use_scheduler:
            from torch.optim.lr_scheduler import LambdaLR
            scheduler = instantiate_from_config(self.scheduler_config)
            return [opt], [{"scheduler": LambdaLR(opt, lr_lambda=scheduler.schedule), "interval": "step", "frequency": 1}]
        return opt

    # ---- conditioning / first stage -------------------------------------------------------------------
    def get_learned_conditioning(self, c):
        """ddpm.py:559-570"""
        if self.cond_stage_model is None:
            return c
        if self.cond_stage_forward is None:
            if hasattr(self.cond_stage_model, "encode") and callable(self.cond_stage_model.encode):
                c = self.cond_stage_model.encode(c)
                if hasattr(c, "mode"):
                    c = c.mode()
            else:
                c = self.cond_stage_model(c)
        else:
            c = getattr(self.cond_stage_model, self.cond_stage_forward)(c)
        return c

    def get_first_stage_encoding(self, encoder_posterior):
        """ddpm.py:550-557"""
        z = encoder_posterior.sample() if hasattr(encoder_posterior, "sample") else encoder_posterior
        return self.scale_factor * z

    @torch.no_grad()
    def encode_first_stage(self, x):
        return self.first_stage_model.encode(x)

    @torch.no_grad()
    def decode_first_stage(self, z, predict_cids=False, force_not_quantize=False):
        """ddpm.py:713-771 (no split_input_params tiling, which EALDM's configs never set)."""
        if predict_cids:                                   # ddpm.py:715-719
            if z.dim() == 4:
                z = torch.argmax(z.exp(), dim=1).long()
            z = self.first_stage_model.quantize.get_codebook_entry(z, shape=None)
            z = z.permute(0, 3, 1, 2).contiguous()
        z = 1. / self.scale_factor * z
        if hasattr(self.first_stage_model, "quantize") and not isinstance(self.first_stage_model, IdentityFirstStage):
            # VQModelInterface (ddpm.py:768-769): callers may ask for the un-quantised latent
            return self.first_stage_model.decode(z, force_not_quantize=predict_cids or force_not_quantize)
        return self.first_stage_model.decode(z)

    # ---- data plumbing of the trainer (ddpm.py:330-343, 662-711, 873-876) ----------------------------------------
    def _get_input_base(self, batch, k):
        x = batch[k]
        if len(x.shape) == 3:
            x = x[..., None]
        elif len(x.shape) == 5:
            x = x.squeeze(0)
        return x.to(memory_format=torch.contiguous_format).float()

    @torch.no_grad()
    def get_input(self, batch, k, return_first_stage_outputs=False, force_c_encode=False, cond_key=None,
                  return_original_cond=False, bs=None):
        x = self._get_input_base(batch, k)
        if bs is not None:
            x = x[:bs]
        x = x.to(self.device)
        z = self.get_first_stage_encoding(self.encode_first_stage(x)).detach()
        c, xc = None, None
        if self.model.conditioning_key is not None:
            cond_key = self.cond_stage_key if cond_key is None else cond_key
            if cond_key != self.first_stage_key:
                if cond_key in ("caption", "coordinates_bbox", "mixed"):
                    xc = batch[cond_key]
                elif cond_key == "class_label":
                    xc = batch
                else:
                    xc = self._get_input_base(batch, cond_key).to(self.device)
            else:
                xc = x
            if not self.cond_stage_trainable or force_c_encode:
                c = self.get_learned_conditioning(xc if isinstance(xc, (dict, list)) else xc.to(self.device))
            else:
                c = xc
            if bs is not None:
                c = c[:bs]
        out = [z, c]
        if return_first_stage_outputs:
            out.extend([x, self.decode_first_stage(z)])
        if return_original_cond:
            out.append(xc)
        return out

    def shared_step(self, batch, **kwargs):
        x, c = self.get_input(batch, self.first_stage_key)
        return self(x, c)

    def training_step(self, batch, batch_idx=0):
        """ddpm.py:345-358 without the Lightning logging calls."""
        loss, _ = self.shared_step(batch)
        return loss

    # ---- denoiser ---------------------------------------------------------------------------------------
    def apply_model(self, x_noisy, t, cond, return_ids=False):
        """ddpm.py:912-921,1011-1016"""
        if not isinstance(cond, dict):
            if not isinstance(cond, list):
                cond = [cond]
            key = "c_concat" if self.model.conditioning_key == "concat" else "c_crossattn"
            cond = {key: cond}
        if self.model.conditioning_key is None:
            cond = {}
        x_recon = self.model(x_noisy, t, **cond)
        if isinstance(x_recon, tuple) and not return_ids:
            return x_recon[0]
        return x_recon

    def q_sample(self, x_start, t, noise=None):
        """ddpm.py:276-279 as one kernel (bit-exact)."""
        noise = torch.randn_like(x_start) if noise is None else noise
        return ops.q_sample(x_start.float().contiguous(), noise.float().contiguous(), t.contiguous(),
                            self.sqrt_alphas_cumprod, self.sqrt_one_minus_alphas_cumprod)

    def forward(self, x, c, *args, **kwargs):
        """ddpm.py:878-900.  With a trainable conditioner `c` is the raw `batch['mixed']` list: the NEGATIVE conditioning
        is the same list with the frames replaced by its last entry (the negative frames) and that entry set to None
        (ddpm.py:886-888), both go through the conditioner and the UNet sees cat([c_neg, c]).  With a frozen conditioner
        `c` is already encoded (get_input) -- for the guided loss the caller passes cat([c_neg, c]) (2B rows)."""
        t = torch.randint(0, self.num_timesteps, (x.shape[0],), device=self.device).long()
        if self.model.conditioning_key is not None:
            assert c is not None
            if self.cond_stage_trainable:
                if self.unconditional_guidance_scale != 1.:
                    c_neg = copy.copy(c) if isinstance(c, list) else copy.deepcopy(c)
                    c_neg[0] = c_neg[-1]
                    c_neg[-1] = None
                    c_neg = self.get_learned_conditioning(c_neg)
                    c = self.get_learned_conditioning(c)
                    c = torch.cat([c_neg, c])
                else:
                    c = self.get_learned_conditioning(c)
        return self.p_losses(x, c, t, *args, **kwargs)

    def p_losses(self, x_start, cond, t, noise=None):
        """ddpm.py:1036-1078: the UNet always sees the doubled batch; the guided eps is regressed on noise."""
        noise = torch.randn_like(x_start) if noise is None else noise
        x_noisy = self.q_sample(x_start=x_start, t=t, noise=noise)
        s = self.unconditional_guidance_scale
        target = noise if self.parameterization == "eps" else x_start
        if s != 1.:
            e_u, e_c = self.apply_model(torch.cat([x_noisy] * 2), torch.cat([t] * 2), cond).chunk(2)
            loss_simple = _CfgMse.apply(e_u, e_c, target, s)
        else:
            loss_simple = _CfgMse.apply(None, self.apply_model(x_noisy, t, cond), target, 1.0)
        prefix = "train" if self.training else "val"
        loss_dict = {f"{prefix}/loss_simple": loss_simple.mean()}
        logvar_t = self.logvar.to(self.device)[t]
        loss = loss_simple / torch.exp(logvar_t) + logvar_t
        if self.learn_logvar:
            loss_dict[f"{prefix}/loss_gamma"] = loss.mean()
            loss_dict["logvar"] = self.logvar.data.mean()
        loss = self.l_simple_weight * loss.mean()
        loss_vlb = (self.lvlb_weights[t] * loss_simple).mean()
        loss_dict[f"{prefix}/loss_vlb"] = loss_vlb
        loss = loss + self.original_elbo_weight * loss_vlb
        loss_dict[f"{prefix}/loss"] = loss
        return loss, loss_dict

    # ---- DDPM ancestral sampling (ddpm.py:1081-1266) ----------------------------------------------------------
    @torch.no_grad()
    def p_sample(self, x, c, t, clip_denoised=False, repeat_noise=False, return_codebook_ids=False,
                 quantize_denoised=False, return_x0=False, temperature=1., noise_dropout=0., score_corrector=None,
                 corrector_kwargs=None, noise=None):
        """ddpm.py:1111-1140: UNet eps, then predict_start_from_noise + q_posterior + the noise term as ONE kernel
        (`ealdm_ddpm_step`).  `noise` (not in the reference) replaces the torch.randn draw, for tests."""
        if return_codebook_ids or quantize_denoised or score_corrector is not None or noise_dropout > 0. or repeat_noise:
            raise NotImplementedError("p_sample options outside the eps / Gaussian-noise path")
        assert self.parameterization == "eps"
        eps = self.apply_model(x, t, c).contiguous()
        nz = torch.randn(x.shape, device=x.device) if noise is None else noise
        bufs = (self.sqrt_recip_alphas_cumprod, self.sqrt_recipm1_alphas_cumprod, self.posterior_mean_coef1,
                self.posterior_mean_coef2, self.posterior_log_variance_clipped)
        return ops.ddpm_step(x.float().contiguous(), eps, nz.float().contiguous(), t.contiguous(), bufs,
                             clip_denoised=clip_denoised, temperature=temperature, want_x0=return_x0)

    @torch.no_grad()
    def p_sample_loop(self, cond, shape, return_intermediates=False, x_T=None, verbose=True, callback=None,
                      timesteps=None, quantize_denoised=False, mask=None, x0=None, img_callback=None, start_T=None,
                      log_every_t=None, noises=None):
        """ddpm.py:1190-1247.  `noises` (not in the reference): per-iteration noise tensors for tests."""
        log_every_t = log_every_t or self.log_every_t
        device = self.betas.device
        revalidate_packed(self)
        b = shape[0]
        img = torch.randn(shape, device=device) if x_T is None else x_T
        intermediates = [img]
        timesteps = self.num_timesteps if timesteps is None else timesteps
        if start_T is not None:
            timesteps = min(timesteps, start_T)
        if mask is not None:
            assert x0 is not None and x0.shape[2:3] == mask.shape[2:3]
        for k, i in enumerate(reversed(range(0, timesteps))):
            ts = torch.full((b,), i, device=device, dtype=torch.long)
            img = self.p_sample(img, cond, ts, clip_denoised=self.clip_denoised, quantize_denoised=quantize_denoised,
                                noise=None if noises is None else noises[k])
            if mask is not None:
                img = self.q_sample(x0, ts) * mask + (1. - mask) * img
            if i % log_every_t == 0 or i == timesteps - 1:
                intermediates.append(img)
            if callback:
                callback(i)
            if img_callback:
                img_callback(img, i)
        return (img, intermediates) if return_intermediates else img

    @torch.no_grad()
    def sample(self, cond, batch_size=16, return_intermediates=False, x_T=None, verbose=True, timesteps=None,
               quantize_denoised=False, mask=None, x0=None, shape=None, **kwargs):
        """ddpm.py:1249-1265"""
        if shape is None:
            shape = (batch_size, self.channels, self.image_size, self.image_size)
        if cond is not None and not isinstance(cond, dict):
            cond = [c[:batch_size] for c in cond] if isinstance(cond, list) else cond[:batch_size]
        return self.p_sample_loop(cond, shape, return_intermediates=return_intermediates, x_T=x_T, verbose=verbose,
                                  timesteps=timesteps, quantize_denoised=quantize_denoised, mask=mask, x0=x0)

    @torch.no_grad()
    def sample_log(self, cond, batch_size, ddim, ddim_steps, **kwargs):
        """ddpm.py:1267-1285: DDIM with the model's guidance scale (or ancestral sampling); `cond` is cat([c_neg, c])."""
        if not ddim:
            samples, intermediates = self.sample(cond=cond, batch_size=batch_size, return_intermediates=True, **kwargs)
            return samples, intermediates
        sampler = DDIMSampler(self)
        shape = (self.channels, self.image_size, self.image_size)
        if self.unconditional_guidance_scale != 1. and cond is not None:
            c_neg, c = cond.chunk(2)
            return sampler.sample(ddim_steps, batch_size, shape, c, verbose=False,
                                  unconditional_guidance_scale=self.unconditional_guidance_scale,
                                  unconditional_conditioning=c_neg, **kwargs)
        return sampler.sample(ddim_steps, batch_size, shape, cond, verbose=False, **kwargs)
