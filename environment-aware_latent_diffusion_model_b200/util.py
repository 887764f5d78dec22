"""Plugin boundary and host-side schedule arithmetic.

`instantiate_from_config` is the reference's plugin mechanism (ldm/util.py:78-93): a config is
`{"target": "pkg.module.Class", "params": {...}}`; swapping the B200 modules in is a YAML edit of the
`target:` strings.  The schedule helpers restate
ldm/modules/diffusionmodules/util.py:21-74 and must stay bit-exact with it (tests/test_host_cpu.py).
"""
from __future__ import annotations

import importlib

import numpy as np
import torch


def get_obj_from_str(string: str, reload: bool = False):
    module, cls = string.rsplit(".", 1)
    mod = importlib.import_module(module, package=None)
    if reload:
        importlib.reload(mod)
    return getattr(mod, cls)


def instantiate_from_config(config):
    if "target" not in config:
        if config in ("__is_first_stage__", "__is_unconditional__"):
            return None
        raise KeyError("Expected key `target` to instantiate.")
    return get_obj_from_str(config["target"])(**config.get("params", dict()))


def count_params(model, verbose=False):
    total = sum(p.numel() for p in model.parameters())
    if verbose:
        print(f"{model.__class__.__name__} has {total * 1.e-6:.2f} M params.")
    return total


def make_beta_schedule(schedule, n_timestep, linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3):
    """util.py:21-43.  float64 throughout; 'linear' is linear in sqrt(beta)."""
    if schedule == "linear":
        betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64) ** 2
    elif schedule == "cosine":
        ts = torch.arange(n_timestep + 1, dtype=torch.float64) / n_timestep + cosine_s
        alphas = torch.cos(ts / (1 + cosine_s) * np.pi / 2).pow(2)
        alphas = alphas / alphas[0]
        betas = np.clip(1 - alphas[1:] / alphas[:-1], a_min=0, a_max=0.999)
    elif schedule == "sqrt_linear":
        betas = torch.linspace(linear_start, linear_end, n_timestep, dtype=torch.float64)
    elif schedule == "sqrt":
        betas = torch.linspace(linear_start, linear_end, n_timestep, dtype=torch.float64) ** 0.5
    else:
        raise ValueError(f"schedule '{schedule}' unknown.")
    return betas.numpy()


def make_ddim_timesteps(ddim_discr_method, num_ddim_timesteps, num_ddpm_timesteps, verbose=True):
    """util.py:46-60: uniform stride c = T // S, shifted by +1."""
    if ddim_discr_method == "uniform":
        c = num_ddpm_timesteps // num_ddim_timesteps
        ddim_timesteps = np.asarray(list(range(0, num_ddpm_timesteps, c)))
    elif ddim_discr_method == "quad":
        ddim_timesteps = ((np.linspace(0, np.sqrt(num_ddpm_timesteps * .8), num_ddim_timesteps)) ** 2).astype(int)
    else:
        raise NotImplementedError(f'There is no ddim discretization method called "{ddim_discr_method}"')
    steps_out = ddim_timesteps + 1
    if verbose:
        print(f"Selected timesteps for ddim sampler: {steps_out}")
    return steps_out


def make_ddim_sampling_parameters(alphacums, ddim_timesteps, eta, verbose=True):
    """util.py:63-74.  `alphacums` is a CPU float32 tensor; operand types are kept exactly as in the
    reference (torch f32 / numpy f64 mix) so every rounding point is the same."""
    alphas = alphacums[ddim_timesteps]
    alphas_prev = np.asarray([alphacums[0]] + alphacums[ddim_timesteps[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    if verbose:
        print(f"Selected alphas for ddim sampler: a_t: {alphas}; a_(t-1): {alphas_prev}")
        print(f"For the chosen value of eta, which is {eta}, "
              f"this results in the following sigma_t schedule for ddim sampler {sigmas}")
    return sigmas, alphas, alphas_prev


def extract_into_tensor(a, t, x_shape):
    """util.py:96-99"""
    b = t.shape[0]
    return a.gather(-1, t).reshape(b, *((1,) * (len(x_shape) - 1)))


def noise_like(shape, device, repeat=False):
    """util.py:264-267"""
    if repeat:
        return torch.randn((1, *shape[1:]), device=device).repeat(shape[0], *((1,) * (len(shape) - 1)))
    return torch.randn(shape, device=device)


def revalidate_packed(model) -> None:
    """Called by the samplers once per `sample()`: every sub-module that caches kernel-layout weights re-checks them
    against its parameters (catches in-place writes through `.data`, e.g. the reference's LitEma.copy_to)."""
    mods = model.modules() if hasattr(model, "modules") else ()
    for m in mods:
        fn = getattr(m, "revalidate_packed", None)
        if callable(fn):
            fn()


def _hosted_unet(model):
    unet = getattr(getattr(model, "model", None), "diffusion_model", None)
    return unet if hasattr(unet, "sampling_scope") else None


def sampling_scope(model):
    """Context manager around one denoising loop: the hosted UNet projects a loop-invariant context once
    (unet.UNetModel.sampling_scope); a no-op for any other model."""
    import contextlib
    unet = _hosted_unet(model)
    return unet.sampling_scope() if unet is not None else contextlib.nullcontext()


class GuidancePair:
    """`apply_model(torch.cat([x] * 2), torch.cat([t] * 2), torch.cat([uc, c]))` of the reference's guided step
    (ddim.py:176-179, plms.py:184-187): the concatenated conditioning is built once per loop while `uc` and `c` stay
    the same tensors (so that the UNet recognises it), and the UNet is told that both halves share x and t."""

    def __init__(self, model):
        self.model, self.unet = model, _hosted_unet(model)
        self._key, self._held, self._c_in = None, (None, None), None

    def reset(self):
        """Drop the cached concatenation (and the references to `uc` / `c` that keep it valid)."""
        self._key, self._held, self._c_in = None, (None, None), None

    def __call__(self, x, t, uc, c):
        import contextlib
        if isinstance(c, torch.Tensor) and isinstance(uc, torch.Tensor):
            key = (id(uc), uc._version, id(c), c._version)
            if key != self._key or self._held[0] is not uc or self._held[1] is not c:
                self._key, self._held, self._c_in = key, (uc, c), torch.cat([uc, c])
            c_in = self._c_in
        else:
            c_in = torch.cat([uc, c])
        with (self.unet.cfg_pair() if self.unet is not None else contextlib.nullcontext()):
            return self.model.apply_model(torch.cat([x] * 2), torch.cat([t] * 2), c_in).chunk(2)
