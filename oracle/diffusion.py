"""Oracle: schedules, DDIM sampling, q_sample and p_losses (TEST INFRASTRUCTURE, see oracle/__init__.py).

The schedule arithmetic reproduces the reference's exact sequence of float64 numpy operations and
float32 rounding points, because the north star requires the DDIM timestep schedule and its tables to
be bit-exact."""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch


# ---- ldm/modules/diffusionmodules/util.py:21-43 ------------------------------------------------------
def make_beta_schedule(n_timestep: int = 1000, linear_start: float = 1e-4, linear_end: float = 2e-2) -> np.ndarray:
    """'linear' schedule only (the one every EALDM config uses): linspace in sqrt space, float64."""
    return (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64) ** 2).numpy()


# ---- ldm/models/diffusion/ddpm.py:119-171 ------------------------------------------------------------
def register_schedule(timesteps: int = 1000, linear_start: float = 1e-4, linear_end: float = 2e-2,
                      v_posterior: float = 0.0) -> Dict[str, torch.Tensor]:
    """The float32 buffers DDPM.register_schedule creates (float64 numpy math, then torch.tensor(..., f32))."""
    betas = make_beta_schedule(timesteps, linear_start, linear_end)
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    f32 = lambda a: torch.tensor(a, dtype=torch.float32)  # noqa: E731
    post_var = (1 - v_posterior) * betas * (1.0 - ac_prev) / (1.0 - ac) + v_posterior * betas
    buf = {
        "betas": f32(betas),
        "alphas_cumprod": f32(ac),
        "alphas_cumprod_prev": f32(ac_prev),
        "sqrt_alphas_cumprod": f32(np.sqrt(ac)),
        "sqrt_one_minus_alphas_cumprod": f32(np.sqrt(1.0 - ac)),
        "log_one_minus_alphas_cumprod": f32(np.log(1.0 - ac)),
        "sqrt_recip_alphas_cumprod": f32(np.sqrt(1.0 / ac)),
        "sqrt_recipm1_alphas_cumprod": f32(np.sqrt(1.0 / ac - 1)),
        "posterior_variance": f32(post_var),
        "posterior_log_variance_clipped": f32(np.log(np.maximum(post_var, 1e-20))),
        "posterior_mean_coef1": f32(betas * np.sqrt(ac_prev) / (1.0 - ac)),
        "posterior_mean_coef2": f32((1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac)),
    }
    # eps-parameterisation weights, ddpm.py:160-170
    lvlb = buf["betas"] ** 2 / (2 * buf["posterior_variance"] * f32(alphas) * (1 - buf["alphas_cumprod"]))
    lvlb[0] = lvlb[1]
    buf["lvlb_weights"] = lvlb
    return buf


# ---- util.py:46-60 -----------------------------------------------------------------------------------
def make_ddim_timesteps(num_ddim_timesteps: int, num_ddpm_timesteps: int = 1000) -> np.ndarray:
    c = num_ddpm_timesteps // num_ddim_timesteps
    return np.asarray(list(range(0, num_ddpm_timesteps, c))) + 1


# ---- util.py:63-74 + ddim.py:24-53 ---------------------------------------------------------------------
def make_ddim_schedule(alphas_cumprod_f32: torch.Tensor, S: int, eta: float) -> Dict[str, object]:
    """Tables DDIMSampler.make_schedule registers.  Dtypes follow what the reference produces with
    torch 2.x / numpy 2.x: ddim_alphas torch f32 (indexing a CPU f32 tensor with a
    numpy int array), ddim_alphas_prev numpy f64 holding f32-exact values, ddim_sigmas torch f64."""
    ts = make_ddim_timesteps(S, alphas_cumprod_f32.shape[0])
    alphacums = alphas_cumprod_f32.cpu()
    alphas = alphacums[ts]
    alphas_prev = np.asarray([alphacums[0]] + alphacums[ts[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    return {
        "ddim_timesteps": ts,
        "ddim_alphas": alphas,
        "ddim_alphas_prev": alphas_prev,
        "ddim_sigmas": sigmas,
        "ddim_sqrt_one_minus_alphas": np.sqrt(1.0 - alphas),
    }


def ddim_step_scalars(sched: Dict[str, object], index: int) -> Tuple[float, float, float, float]:
    """ddim.py:189-192: torch.full((b,1,1,1), table[index]) -> float32 values a_t, a_prev, sigma_t,
    sqrt(1-a_t), returned as float32 tensors of shape [] so later arithmetic stays in fp32."""
    def as_f32(v):
        return torch.full((), float(v), dtype=torch.float32)
    return (as_f32(sched["ddim_alphas"][index]), as_f32(sched["ddim_alphas_prev"][index]),
            as_f32(sched["ddim_sigmas"][index]), as_f32(sched["ddim_sqrt_one_minus_alphas"][index]))


# ---- ddim.py:165-203 -----------------------------------------------------------------------------------
def p_sample_ddim(apply_model: Callable, x: torch.Tensor, c, t: torch.Tensor, sched, index: int,
                  noise: torch.Tensor, temperature: float = 1.0, ugs: float = 1.0, uc=None):
    if uc is None or ugs == 1.0:
        e_t = apply_model(x, t, c)
    else:
        x_in, t_in = torch.cat([x] * 2), torch.cat([t] * 2)
        c_in = torch.cat([uc, c])
        e_u, e_c = apply_model(x_in, t_in, c_in).chunk(2)
        e_t = e_u + ugs * (e_c - e_u)
    a_t, a_prev, sigma_t, s1 = ddim_step_scalars(sched, index)
    pred_x0 = (x - s1 * e_t) / a_t.sqrt()
    dir_xt = (1.0 - a_prev - sigma_t ** 2).sqrt() * e_t
    nz = sigma_t * noise * temperature
    x_prev = a_prev.sqrt() * pred_x0 + dir_xt + nz
    return x_prev, pred_x0, e_t


# ---- ddim.py:113-162 -----------------------------------------------------------------------------------
def ddim_sample(apply_model: Callable, alphas_cumprod_f32: torch.Tensor, S: int, x_T: torch.Tensor, cond=None,
                eta: float = 0.0, ugs: float = 1.0, uc=None, noises: Optional[List[torch.Tensor]] = None,
                temperature: float = 1.0):
    """Returns (x_0, trace) with trace = per-step dicts {'t','e_t','x_prev','pred_x0'}.
    `noises[i]` replaces the torch.randn draw of step i (the reference always draws, ddim.py:200)."""
    sched = make_ddim_schedule(alphas_cumprod_f32, S, eta)
    ts = sched["ddim_timesteps"]
    img = x_T
    b = x_T.shape[0]
    trace = []
    total = ts.shape[0]
    for i, step in enumerate(np.flip(ts)):
        index = total - i - 1
        t = torch.full((b,), int(step), dtype=torch.long)
        nz = noises[i] if noises is not None else torch.zeros_like(img)
        img, pred_x0, e_t = p_sample_ddim(apply_model, img, cond, t, sched, index, nz, temperature, ugs, uc)
        trace.append({"t": int(step), "e_t": e_t, "x_prev": img, "pred_x0": pred_x0})
    return img, trace


# ---- ddpm.py:1081-1140, :218-231 --------------------------------------------------------------------------
def p_sample_ddpm(apply_model: Callable, buf: Dict[str, torch.Tensor], x: torch.Tensor, c, t: torch.Tensor,
                  noise: torch.Tensor, clip_denoised: bool = False, temperature: float = 1.0):
    """LatentDiffusion.p_sample (eps parameterisation): returns (x_prev, x0)."""
    b = x.shape[0]
    ex = lambda a: a[t].reshape(b, 1, 1, 1)  # noqa: E731  (extract_into_tensor, util.py:96-99)
    eps = apply_model(x, t, c)
    x0 = ex(buf["sqrt_recip_alphas_cumprod"]) * x - ex(buf["sqrt_recipm1_alphas_cumprod"]) * eps
    if clip_denoised:
        x0 = x0.clamp(-1.0, 1.0)
    mean = ex(buf["posterior_mean_coef1"]) * x0 + ex(buf["posterior_mean_coef2"]) * x
    logvar = ex(buf["posterior_log_variance_clipped"])
    nonzero_mask = (1 - (t == 0).float()).reshape(b, 1, 1, 1)
    return mean + nonzero_mask * (0.5 * logvar).exp() * (noise * temperature), x0


# ---- plms.py:120-236 -------------------------------------------------------------------------------------
def plms_sample(apply_model: Callable, alphas_cumprod_f32: torch.Tensor, S: int, x_T: torch.Tensor, cond=None,
                ugs: float = 1.0, uc=None):
    """PLMSSampler.plms_sampling + p_sample_plms (eta = 0): pseudo improved Euler on the first step, then
    Adams-Bashforth of order 2, 3, 4 on the eps history.  Returns (x_0, trace of per-step x_prev / pred_x0 / e_t)."""
    sched = make_ddim_schedule(alphas_cumprod_f32, S, 0.0)
    ts = np.flip(sched["ddim_timesteps"])
    total = ts.shape[0]
    b = x_T.shape[0]
    img, old_eps, trace = x_T, [], []

    def model_eps(x, t):
        if uc is None or ugs == 1.0:
            return apply_model(x, t, cond)
        e_u, e_c = apply_model(torch.cat([x] * 2), torch.cat([t] * 2), torch.cat([uc, cond])).chunk(2)
        return e_u + ugs * (e_c - e_u)

    def x_prev_and_x0(x, e, index):
        a_t, a_prev, sigma_t, s1 = ddim_step_scalars(sched, index)
        pred_x0 = (x - s1 * e) / a_t.sqrt()
        dir_xt = (1.0 - a_prev - sigma_t ** 2).sqrt() * e
        return a_prev.sqrt() * pred_x0 + dir_xt, pred_x0

    for i, step in enumerate(ts):
        index = total - i - 1
        t = torch.full((b,), int(step), dtype=torch.long)
        t_next = torch.full((b,), int(ts[min(i + 1, total - 1)]), dtype=torch.long)
        e_t = model_eps(img, t)
        if len(old_eps) == 0:
            x_prev, _ = x_prev_and_x0(img, e_t, index)
            e_prime = (e_t + model_eps(x_prev, t_next)) / 2
        elif len(old_eps) == 1:
            e_prime = (3 * e_t - old_eps[-1]) / 2
        elif len(old_eps) == 2:
            e_prime = (23 * e_t - 16 * old_eps[-1] + 5 * old_eps[-2]) / 12
        else:
            e_prime = (55 * e_t - 59 * old_eps[-1] + 37 * old_eps[-2] - 9 * old_eps[-3]) / 24
        img, pred_x0 = x_prev_and_x0(img, e_prime, index)
        old_eps.append(e_t)
        if len(old_eps) >= 4:
            old_eps.pop(0)
        trace.append({"t": int(step), "e_t": e_t, "x_prev": img, "pred_x0": pred_x0})
    return img, trace


# ---- ddpm.py:276-279, util.py:96-99 ---------------------------------------------------------------------
def q_sample(buf: Dict[str, torch.Tensor], x0: torch.Tensor, t: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
    b = t.shape[0]
    shape = (b,) + (1,) * (x0.dim() - 1)
    return (buf["sqrt_alphas_cumprod"].gather(-1, t).reshape(shape) * x0 +
            buf["sqrt_one_minus_alphas_cumprod"].gather(-1, t).reshape(shape) * noise)


# ---- ddpm.py:1036-1078 ------------------------------------------------------------------------------------
def p_losses(apply_model: Callable, buf: Dict[str, torch.Tensor], x0: torch.Tensor, cond, t: torch.Tensor,
             noise: torch.Tensor, ugs: float = 2.0, l_simple_weight: float = 1.0,
             original_elbo_weight: float = 0.0, logvar: Optional[torch.Tensor] = None):
    """EALDM's p_losses: the UNet always runs on the doubled batch and the CFG-combined eps (scale 2,
    ddpm.py:442) is regressed on the noise.  `cond` is the already concatenated [c_neg; c] (2B rows)."""
    x_noisy = q_sample(buf, x0, t, noise)
    if ugs != 1.0:
        e_u, e_c = apply_model(torch.cat([x_noisy] * 2), torch.cat([t] * 2), cond).chunk(2)
        out = e_u + ugs * (e_c - e_u)
    else:
        out = apply_model(x_noisy, t, cond)
    loss_simple = torch.nn.functional.mse_loss(noise, out, reduction="none").mean([1, 2, 3])
    lv = torch.zeros(buf["betas"].shape[0]) if logvar is None else logvar
    logvar_t = lv[t]
    loss = l_simple_weight * (loss_simple / torch.exp(logvar_t) + logvar_t).mean()
    loss_vlb = (buf["lvlb_weights"][t] * loss_simple).mean()
    loss = loss + original_elbo_weight * loss_vlb
    return loss, {"loss_simple": loss_simple.mean(), "loss_vlb": loss_vlb, "loss": loss,
                  "loss_simple_per_sample": loss_simple, "model_output": out}
