"""Drop-in `PLMSSampler` (reference: ldm/models/diffusion/plms.py:11-236) -- SURVEY.md section 8f rank 4.

Same constructor / `sample(...)` contract, the same timestep walk (`ts_next`, plms.py:140-141), the same eps history
of three entries and the same RNG draw order (x_T, then one `randn` per x_prev evaluation even though sigma = 0,
plms.py:204).  The UNet is the same hot path as in DDIM; the sampler-specific arithmetic is two kernels per step:
`ealdm_plms_eps` (CFG combine + Adams-Bashforth combination, bit-exact fp32 operation order) and `ealdm_ddim_step`
(x0 prediction + direction)."""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .ddim import DDIMSampler
from .util import revalidate_packed, sampling_scope


class PLMSSampler(DDIMSampler):
    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0., verbose=True):
        if ddim_eta != 0:
            raise ValueError("ddim_eta must be 0 for PLMS")
        super().make_schedule(ddim_num_steps, ddim_discretize=ddim_discretize, ddim_eta=ddim_eta, verbose=verbose)

    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None,
               img_callback=None, quantize_x0=False, eta=0., mask=None, x0=None, temperature=1.,
               noise_dropout=0., score_corrector=None, corrector_kwargs=None, verbose=True, x_T=None,
               log_every_t=100, unconditional_guidance_scale=1., unconditional_conditioning=None, **kwargs):
        if quantize_x0 or score_corrector is not None or noise_dropout > 0.:
            raise NotImplementedError("PLMS with quantize_x0 / score_corrector / noise_dropout is not built")
        if conditioning is not None:
            cbs = (conditioning[list(conditioning.keys())[0]].shape[0] if isinstance(conditioning, dict)
                   else conditioning.shape[0])
            if cbs != batch_size:
                print(f"Warning: Got {cbs} conditionings but batch-size is {batch_size}")
        self.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=verbose)
        revalidate_packed(self.model)
        C, H, W = shape
        return self.plms_sampling(conditioning, (batch_size, C, H, W), callback=callback, img_callback=img_callback,
                                  mask=mask, x0=x0, temperature=temperature, x_T=x_T, log_every_t=log_every_t,
                                  unconditional_guidance_scale=unconditional_guidance_scale,
                                  unconditional_conditioning=unconditional_conditioning)

    @torch.no_grad()
    def plms_sampling(self, cond, shape, x_T=None, callback=None, timesteps=None, mask=None, x0=None,
                      img_callback=None, log_every_t=100, temperature=1., unconditional_guidance_scale=1.,
                      unconditional_conditioning=None):
        device = self.model.betas.device
        b = shape[0]
        img = torch.randn(shape, device=device) if x_T is None else x_T
        if timesteps is None:
            timesteps = self.ddim_timesteps
        else:
            subset_end = int(min(timesteps / self.ddim_timesteps.shape[0], 1) * self.ddim_timesteps.shape[0]) - 1
            timesteps = self.ddim_timesteps[:subset_end]
        intermediates = {"x_inter": [img], "pred_x0": [img]}
        time_range = np.flip(timesteps)
        total_steps = timesteps.shape[0]
        ts_all = torch.as_tensor(np.ascontiguousarray(np.asarray(time_range, dtype=np.int64))).to(device)
        ts_all = ts_all[:, None].expand(total_steps, b).contiguous()
        old_eps = []
        self._pair.reset()
        with sampling_scope(self.model):     # the conditioning is loop-invariant: the UNet projects it once
            for i in range(total_steps):
                index = total_steps - i - 1
                ts, ts_next = ts_all[i], ts_all[min(i + 1, total_steps - 1)]
                if mask is not None:
                    assert x0 is not None
                    img_orig = self.model.q_sample(x0, ts)
                    img = img_orig * mask + (1. - mask) * img
                img, pred_x0, e_t = self.p_sample_plms(img, cond, ts, index=index, temperature=temperature,
                                                       unconditional_guidance_scale=unconditional_guidance_scale,
                                                       unconditional_conditioning=unconditional_conditioning,
                                                       old_eps=old_eps, t_next=ts_next)
                old_eps.append(e_t)
                if len(old_eps) >= 4:
                    old_eps.pop(0)
                if callback:
                    callback(i)
                if img_callback:
                    img_callback(pred_x0, i)
                if index % log_every_t == 0 or index == total_steps - 1:
                    intermediates["x_inter"].append(img)
                    intermediates["pred_x0"].append(pred_x0)
        return img, intermediates

    def _model_eps(self, x, c, t, ugs, uc):
        """(e_uncond or None, e_cond): plms.py:181-189 without the combine (done inside ealdm_plms_eps)."""
        if uc is None or ugs == 1.:
            return None, self.model.apply_model(x, t, c).contiguous()
        e_u, e_c = self._pair(x, t, uc, c)
        return e_u.contiguous(), e_c.contiguous()

    def _x_prev(self, x, e, index, temperature):
        """get_x_prev_and_pred_x0 (plms.py:200-216): sigma = 0, but the noise is still drawn (RNG stream parity)."""
        noise = torch.randn(x.shape, device=x.device) if self.noise_fn is None else self.noise_fn(x.shape, x.device)
        return ops.ddim_step(x.contiguous(), e, noise=noise, temperature=float(temperature), **self.step_scalars(index))

    @torch.no_grad()
    def p_sample_plms(self, x, c, t, index, temperature=1., unconditional_guidance_scale=1.,
                      unconditional_conditioning=None, old_eps=None, t_next=None, **unused):
        ugs, uc = float(unconditional_guidance_scale), unconditional_conditioning
        e_u, e_c = self._model_eps(x, c, t, ugs, uc)
        n_old = len(old_eps)
        if n_old == 0:
            # pseudo improved Euler (2nd order): a second UNet evaluation at the next timestep
            e_t, _ = ops.plms_eps(e_c, e_uncond=e_u, cfg_scale=ugs, mode=0)
            x_prev, _ = self._x_prev(x, e_t, index, temperature)
            e_u2, e_c2 = self._model_eps(x_prev, c, t_next, ugs, uc)
            e_next, _ = ops.plms_eps(e_c2, e_uncond=e_u2, cfg_scale=ugs, mode=0)
            _, e_prime = ops.plms_eps(e_t, old=(e_next,), mode=1)
        else:
            hist = tuple(old_eps[::-1][:3])     # old_eps[-1], old_eps[-2], old_eps[-3]
            e_t, e_prime = ops.plms_eps(e_c, e_uncond=e_u, cfg_scale=ugs, old=hist, mode=min(n_old, 3) + 1)
        x_prev, pred_x0 = self._x_prev(x, e_prime, index, temperature)
        return x_prev, pred_x0, e_t
