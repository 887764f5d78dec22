"""Drop-in `DDIMSampler` (reference: ldm/models/diffusion/ddim.py:11-203).

Same constructor and `sample(...)` contract and the same RNG draw order (x_T, then one `randn` per
step, always drawn -- ddim.py:122,200).  Differences in mechanism, not in results:
  * the schedule tables stay on the host (the reference indexes CUDA tensors, three device->host
    syncs per step, ddim.py:189-192); the per-step fp32 scalars are computed once per `sample()`;
  * the CFG combine, x0 prediction, direction and noise terms are ONE kernel (`ealdm_ddim_step`),
    bit-exact with the reference's chain of ~20 elementwise launches.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .util import GuidancePair, make_ddim_sampling_parameters, make_ddim_timesteps, revalidate_packed, sampling_scope


class DDIMSampler(object):
    def __init__(self, model, schedule="linear", **kwargs):
        super().__init__()
        self.model = model
        self.ddpm_num_timesteps = model.num_timesteps
        self.schedule = schedule
        # test hook: callable(shape, device) -> noise replacing the per-step torch.randn draw
        self.noise_fn = kwargs.get("noise_fn", None)
        self._pair = GuidancePair(model)

    def register_buffer(self, name, attr):
        setattr(self, name, attr)

    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0., verbose=True):
        """ddim.py:24-53.  All tables are host objects with the reference's dtypes."""
        self.ddim_timesteps = make_ddim_timesteps(ddim_discr_method=ddim_discretize,
                                                  num_ddim_timesteps=ddim_num_steps,
                                                  num_ddpm_timesteps=self.ddpm_num_timesteps, verbose=verbose)
        alphas_cumprod = self.model.alphas_cumprod
        assert alphas_cumprod.shape[0] == self.ddpm_num_timesteps, "alphas have to be defined for each timestep"
        ac = alphas_cumprod.detach().to(torch.float32).cpu()
        self.alphas_cumprod = ac
        self.alphas_cumprod_prev = self.model.alphas_cumprod_prev.detach().to(torch.float32).cpu()
        self.betas = self.model.betas.detach().to(torch.float32).cpu()
        self.sqrt_alphas_cumprod = torch.as_tensor(np.sqrt(ac.numpy()))
        self.sqrt_one_minus_alphas_cumprod = torch.as_tensor(np.sqrt(1. - ac.numpy()))
        ddim_sigmas, ddim_alphas, ddim_alphas_prev = make_ddim_sampling_parameters(
            alphacums=ac, ddim_timesteps=self.ddim_timesteps, eta=ddim_eta, verbose=verbose)
        self.ddim_sigmas = ddim_sigmas
        self.ddim_alphas = ddim_alphas
        self.ddim_alphas_prev = ddim_alphas_prev
        self.ddim_sqrt_one_minus_alphas = np.sqrt(1. - ddim_alphas)
        self.ddim_sigmas_for_original_num_steps = ddim_eta * torch.sqrt(
            (1 - self.alphas_cumprod_prev) / (1 - self.alphas_cumprod) *
            (1 - self.alphas_cumprod / self.alphas_cumprod_prev))

    def step_scalars(self, index, use_original_steps=False):
        """The fp32 values the reference materialises with torch.full (ddim.py:189-192) and the fp32
        tensor arithmetic it applies to them (:195-199): sqrt(a_t), sqrt(a_prev), sqrt(1-a_prev-sigma^2)."""
        if use_original_steps:
            alphas, alphas_prev = self.alphas_cumprod, self.alphas_cumprod_prev
            s1, sigmas = self.sqrt_one_minus_alphas_cumprod, self.ddim_sigmas_for_original_num_steps
        else:
            alphas, alphas_prev = self.ddim_alphas, self.ddim_alphas_prev
            s1, sigmas = self.ddim_sqrt_one_minus_alphas, self.ddim_sigmas
        f = lambda v: torch.full((), float(v), dtype=torch.float32)  # noqa: E731
        a_t, a_prev, sigma_t, s1t = f(alphas[index]), f(alphas_prev[index]), f(sigmas[index]), f(s1[index])
        return {
            "sqrt_one_minus_at": float(s1t),
            "sqrt_at": float(a_t.sqrt()),
            "sqrt_a_prev": float(a_prev.sqrt()),
            "dir_coef": float((1. - a_prev - sigma_t ** 2).sqrt()),
            "sigma_t": float(sigma_t),
        }

    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None,
               img_callback=None, quantize_x0=False, eta=0., mask=None, x0=None, temperature=1.,
               noise_dropout=0., score_corrector=None, corrector_kwargs=None, verbose=True, x_T=None,
               log_every_t=100, unconditional_guidance_scale=1., unconditional_conditioning=None, **kwargs):
        if conditioning is not None:
            cbs = (conditioning[list(conditioning.keys())[0]].shape[0] if isinstance(conditioning, dict)
                   else conditioning.shape[0])
            if cbs != batch_size:
                print(f"Warning: Got {cbs} conditionings but batch-size is {batch_size}")
        self.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=verbose)
        revalidate_packed(self.model)
        C, H, W = shape
        size = (batch_size, C, H, W)
        if verbose:
            print(f"Data shape for DDIM sampling is {size}, eta {eta}")
        return self.ddim_sampling(conditioning, size, callback=callback, img_callback=img_callback,
                                  quantize_denoised=quantize_x0, mask=mask, x0=x0, ddim_use_original_steps=False,
                                  noise_dropout=noise_dropout, temperature=temperature,
                                  score_corrector=score_corrector, corrector_kwargs=corrector_kwargs, x_T=x_T,
                                  log_every_t=log_every_t,
                                  unconditional_guidance_scale=unconditional_guidance_scale,
                                  unconditional_conditioning=unconditional_conditioning)

    @torch.no_grad()
    def ddim_sampling(self, cond, shape, x_T=None, ddim_use_original_steps=False, callback=None, timesteps=None,
                      quantize_denoised=False, mask=None, x0=None, img_callback=None, log_every_t=100,
                      temperature=1., noise_dropout=0., score_corrector=None, corrector_kwargs=None,
                      unconditional_guidance_scale=1., unconditional_conditioning=None):
        device = self.model.betas.device
        b = shape[0]
        img = torch.randn(shape, device=device) if x_T is None else x_T
        if timesteps is None:
            timesteps = self.ddpm_num_timesteps if ddim_use_original_steps else self.ddim_timesteps
        elif not ddim_use_original_steps:
            subset_end = int(min(timesteps / self.ddim_timesteps.shape[0], 1) * self.ddim_timesteps.shape[0]) - 1
            timesteps = self.ddim_timesteps[:subset_end]
        intermediates = {"x_inter": [img], "pred_x0": [img]}
        time_range = list(reversed(range(0, timesteps))) if ddim_use_original_steps else np.flip(timesteps)
        total_steps = timesteps if ddim_use_original_steps else timesteps.shape[0]
        # all timestep vectors of the loop in one upload (the reference builds one per step)
        ts_all = torch.as_tensor(np.ascontiguousarray(np.asarray(time_range, dtype=np.int64))).to(device)
        ts_all = ts_all[:, None].expand(total_steps, b).contiguous()
        self._pair.reset()
        with sampling_scope(self.model):   # the conditioning is loop-invariant: the UNet projects it once
            return self._ddim_loop(img, cond, ts_all, total_steps, mask, x0, ddim_use_original_steps, quantize_denoised,
                                   temperature, noise_dropout, score_corrector, corrector_kwargs,
                                   unconditional_guidance_scale, unconditional_conditioning, callback, img_callback,
                                   log_every_t, intermediates)

    def _ddim_loop(self, img, cond, ts_all, total_steps, mask, x0, ddim_use_original_steps, quantize_denoised,
                   temperature, noise_dropout, score_corrector, corrector_kwargs, unconditional_guidance_scale,
                   unconditional_conditioning, callback, img_callback, log_every_t, intermediates):
        for i in range(total_steps):
            index = total_steps - i - 1
            ts = ts_all[i]
            if mask is not None:
                assert x0 is not None
                img_orig = self.model.q_sample(x0, ts)
                img = img_orig * mask + (1. - mask) * img
            img, pred_x0 = self.p_sample_ddim(img, cond, ts, index=index, use_original_steps=ddim_use_original_steps,
                                              quantize_denoised=quantize_denoised, temperature=temperature,
                                              noise_dropout=noise_dropout, score_corrector=score_corrector,
                                              corrector_kwargs=corrector_kwargs,
                                              unconditional_guidance_scale=unconditional_guidance_scale,
                                              unconditional_conditioning=unconditional_conditioning)
            if callback:
                callback(i)
            if img_callback:
                img_callback(pred_x0, i)
            if index % log_every_t == 0 or index == total_steps - 1:
                intermediates["x_inter"].append(img)
                intermediates["pred_x0"].append(pred_x0)
        return img, intermediates

    @torch.no_grad()
    def p_sample_ddim(self, x, c, t, index, repeat_noise=False, use_original_steps=False, quantize_denoised=False,
                      temperature=1., noise_dropout=0., score_corrector=None, corrector_kwargs=None,
                      unconditional_guidance_scale=1., unconditional_conditioning=None, return_eps=False):
        """ddim.py:165-203 with the elementwise tail as one kernel."""
        e_u = None
        if unconditional_conditioning is None or unconditional_guidance_scale == 1.:
            e_c = self.model.apply_model(x, t, c)
        else:
            e_u, e_c = self._pair(x, t, unconditional_conditioning, c)
        sc = self.step_scalars(index, use_original_steps)
        generic = score_corrector is not None or quantize_denoised or noise_dropout > 0. or repeat_noise
        if generic:
            return self._p_sample_generic(x, c, t, e_u, e_c, sc, repeat_noise, quantize_denoised, temperature,
                                          noise_dropout, score_corrector, corrector_kwargs,
                                          unconditional_guidance_scale)
        # always drawn (ddim.py:200), which keeps the Philox stream in step with the reference
        noise = torch.randn(x.shape, device=x.device) if self.noise_fn is None else self.noise_fn(x.shape, x.device)
        out = ops.ddim_step(x.contiguous(), e_c.contiguous(), e_uncond=None if e_u is None else e_u.contiguous(),
                            noise=noise, cfg_scale=float(unconditional_guidance_scale),
                            temperature=float(temperature), want_e=return_eps, **sc)
        return out

    def _p_sample_generic(self, x, c, t, e_u, e_c, sc, repeat_noise, quantize_denoised, temperature, noise_dropout,
                          score_corrector, corrector_kwargs, ugs):
        """Rarely used options (score corrector, VQ re-quantisation of x0, noise dropout): the reference's
        op sequence in torch, since they interleave user callbacks with the update."""
        from .util import noise_like
        e_t = e_c if e_u is None else e_u + ugs * (e_c - e_u)
        if score_corrector is not None:
            assert self.model.parameterization == "eps"
            e_t = score_corrector.modify_score(self.model, e_t, x, t, c, **(corrector_kwargs or {}))
        pred_x0 = (x - sc["sqrt_one_minus_at"] * e_t) / sc["sqrt_at"]
        if quantize_denoised:
            pred_x0, _, *_ = self.model.first_stage_model.quantize(pred_x0)
        dir_xt = sc["dir_coef"] * e_t
        noise = sc["sigma_t"] * noise_like(x.shape, x.device, repeat_noise) * temperature
        if noise_dropout > 0.:
            noise = torch.nn.functional.dropout(noise, p=noise_dropout)
        return sc["sqrt_a_prev"] * pred_x0 + dir_xt + noise, pred_x0
