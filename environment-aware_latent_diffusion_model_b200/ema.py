"""`LitEma` with the reference's checkpoint layout (reference: ldm/modules/ema.py:5-76).

The reference keeps one shadow buffer per trainable parameter of `LatentDiffusion.model`, registered under the
parameter's name with the dots removed (buffer names may not contain '.'), plus `decay` and `num_updates`; they are
saved as `model_ema.*` in every checkpoint and every sampling script runs under `model.ema_scope()`
(ddpm.py:173-186).  This module reproduces that state-dict layout and the update arithmetic (same operation order:
`shadow -= (1 - decay) * (shadow - param)`), batched over all tensors with `torch._foreach_*` instead of a 626-iteration
Python loop.  With `optim.FusedAdamWEMA` the shadows can instead be VIEWS of the optimizer's flat EMA buffer (`bind`):
the fused AdamW+EMA kernel then updates them and this module only provides names, `ema_scope` and checkpoints.

`copy_to` / `restore` write parameters in place, which the packed kernel-layout weights of the UNet cannot see:
both call `invalidate_packed()` on every sub-module that has one.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
from torch import nn


def invalidate_packed_modules(model: nn.Module) -> None:
    for m in model.modules():
        fn = getattr(m, "invalidate_packed", None)
        if callable(fn):
            fn()


class LitEma(nn.Module):
    def __init__(self, model: nn.Module, decay: float = 0.9999, use_num_upates: bool = True):
        super().__init__()
        if decay < 0.0 or decay > 1.0:
            raise ValueError("Decay must be between 0 and 1")
        self.m_name2s_name = {}
        self.register_buffer("decay", torch.tensor(decay, dtype=torch.float32))
        self.register_buffer("num_updates", torch.tensor(0 if use_num_upates else -1, dtype=torch.int))
        for name, p in model.named_parameters():
            if p.requires_grad:
                s_name = name.replace(".", "")
                self.m_name2s_name[name] = s_name
                self.register_buffer(s_name, p.detach().clone())
        self.collected_params: List[torch.Tensor] = []
        self._fused = None          # optim.FusedAdamWEMA once bound
        self._num_updates_host = 0 if use_num_upates else -1   # mirror of `num_updates` (no D2H sync per step)
        self.register_load_state_dict_post_hook(lambda m, keys: m._after_load())

    # ---- pairing of parameters and shadows ---------------------------------------------------------------
    def _pairs(self, model: nn.Module):
        shadows = dict(self.named_buffers())
        ps, ss = [], []
        for name, p in model.named_parameters():
            if p.requires_grad:
                ps.append(p)
                ss.append(shadows[self.m_name2s_name[name]])
            else:
                assert name not in self.m_name2s_name
        return ps, ss

    def _after_load(self):
        self._num_updates_host = int(self.num_updates)
        if self._fused is not None:
            self._fused.num_updates = self._num_updates_host

    def bind(self, fused, model: nn.Module) -> "LitEma":
        """Make every shadow a view of `fused.ema` (optim.FusedAdamWEMA), carrying the current shadow values over."""
        assert fused.ema is not None, "FusedAdamWEMA was built with use_ema=False"
        views = fused.ema_views()
        with torch.no_grad():
            for name, p in model.named_parameters():
                if not p.requires_grad:
                    continue
                s_name = self.m_name2s_name[name]
                v = views[id(p)]
                v.copy_(getattr(self, s_name))
                self._buffers[s_name] = v
        self._fused = fused
        fused.num_updates = self._num_updates_host
        fused.ema_decay = float(self.decay)
        return self

    def state_dict(self, *args, **kwargs):
        if self._fused is not None:       # the fused kernel counts updates on the host
            self._num_updates_host = self._fused.num_updates
        self.num_updates.fill_(self._num_updates_host)
        return super().state_dict(*args, **kwargs)

    # ---- ema.py:25-44 --------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, model: nn.Module):
        if self._fused is not None:
            raise RuntimeError("LitEma is bound to FusedAdamWEMA: its step() already updates the shadows")
        decay = torch.tensor(float(self.decay), dtype=torch.float32)          # host copies: no device sync per step
        if self._num_updates_host >= 0:
            self._num_updates_host += 1
            n = torch.tensor(self._num_updates_host, dtype=torch.int)
            decay = torch.minimum(decay, (1 + n) / (10 + n))                  # int / int -> float32, as the reference
        omd = float(1.0 - decay)                                               # float32 arithmetic, exact as a double
        ps, ss = self._pairs(model)
        diff = torch._foreach_sub(ss, [p.detach() for p in ps])
        torch._foreach_mul_(diff, omd)
        torch._foreach_sub_(ss, diff)

    # ---- ema.py:46-76 --------------------------------------------------------------------------------------
    @torch.no_grad()
    def copy_to(self, model: nn.Module):
        ps, ss = self._pairs(model)
        torch._foreach_copy_([p.data for p in ps], ss)
        if self._fused is not None:
            self._fused.refresh()
        invalidate_packed_modules(model)

    def store(self, parameters: Iterable[nn.Parameter]):
        self.collected_params = [p.detach().clone() for p in parameters]

    @torch.no_grad()
    def restore(self, parameters: Iterable[nn.Parameter], model: Optional[nn.Module] = None):
        ps = list(parameters)
        torch._foreach_copy_([p.data for p in ps], self.collected_params)
        self.collected_params = []
        if self._fused is not None:
            self._fused.refresh()
        if model is not None:
            invalidate_packed_modules(model)
