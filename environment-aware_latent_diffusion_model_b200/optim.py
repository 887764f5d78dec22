"""Fused AdamW + EMA step for the UNet.

Reference: `LatentDiffusion.configure_optimizers` (ldm/models/diffusion/ddpm.py:1409-1431: torch.optim.AdamW over the
UNet parameters) and `LitEma` (ldm/modules/ema.py:25-44: one shadow tensor per parameter, updated by a Python loop after
every step, `ddpm.py:on_train_batch_end`).  Here parameters, gradients (parallel.GradBuckets), both Adam moments, the EMA
shadow and a bf16 copy of the weights are six flat buffers in the same order, and one kernel (`ealdm_adamw_ema_step`)
does the whole update.  The bf16 copy is what the next step's weight packing reads (`UNetEngine._c`), so the separate
fp32 -> bf16 cast of 395 M parameters disappears from the step as well.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict

import torch

from . import _lib as L
from .parallel import GradBuckets


class FusedAdamWEMA:
    def __init__(self, buckets: GradBuckets, lr: float, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, ema_decay: float = 0.9999, use_ema: bool = True,
                 use_num_updates: bool = True):
        self.buckets = buckets
        self.params = buckets.params
        self.unet = getattr(buckets, "unet", None)      # whose packed inference weights every update invalidates
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.ema_decay, self.use_ema = ema_decay, use_ema
        self.step_count = 0
        self.num_updates = 0 if use_num_updates else -1
        total = buckets.flat.numel()
        assert total % 4 == 0
        dev = buckets.flat.device
        with torch.no_grad():
            self.flat_param = torch.zeros(total, dtype=torch.float32, device=dev)
            for p, off in zip(self.params, buckets.offsets):   # parameters become views of the flat buffer
                self.flat_param[off:off + p.numel()].copy_(p.detach().float().reshape(-1))
                p.data = self.flat_param[off:off + p.numel()].view_as(p)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        self.ema = self.flat_param.clone() if use_ema else None
        self.flat_bf16 = self.flat_param.to(torch.bfloat16)
        self._attach_bf16_views()

    def _attach_bf16_views(self):
        for p, off in zip(self.params, self.buckets.offsets):
            p._bf16_view = self.flat_bf16[off:off + p.numel()].view_as(p)

    def _invalidate(self):
        """The UNet caches kernel-layout weights (and CUDA graphs over them) for inference; they are stale now."""
        if self.unet is not None:
            self.unet.invalidate_packed()

    def refresh(self):
        """Call after changing parameters outside `step()` (load_state_dict, manual edits)."""
        self.flat_bf16.copy_(self.flat_param)
        self._invalidate()

    def zero_grad(self):
        self.buckets.zero_()

    @torch.no_grad()
    def step(self, grads_are_sums: bool = False):
        """One update.  `grads_are_sums`: the flat gradient buffer holds the all-reduced SUM over ranks (the division by
        the world size is then folded into this kernel instead of a separate pass over the gradients)."""
        self.step_count += 1
        decay = self.ema_decay
        if self.num_updates >= 0:
            self.num_updates += 1
            decay = min(self.ema_decay, (1 + self.num_updates) / (10 + self.num_updates))
        a = L.AdamWArgs()
        a.param, a.grad = self.flat_param.data_ptr(), self.buckets.flat.data_ptr()
        a.exp_avg, a.exp_avg_sq = self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr()
        a.ema = self.ema.data_ptr() if self.ema is not None else None
        a.param_bf16 = self.flat_bf16.data_ptr()
        a.numel, a.step = self.flat_param.numel(), self.step_count
        a.lr, a.beta1, a.beta2, a.eps, a.weight_decay = self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay
        a.grad_scale = 1.0 / self.buckets.world if grads_are_sums else 1.0
        a.ema_decay = decay
        L.check(L.load().ealdm_adamw_ema_step(C.byref(a), torch.cuda.current_stream().cuda_stream))
        self._invalidate()

    # ---- checkpointing: the per-parameter layout of torch.optim.AdamW.state_dict() ------------------------------
    def state_dict(self) -> dict:
        """Same structure as `torch.optim.AdamW(params).state_dict()` over `self.params` (in this optimizer's
        parameter order): state[i] = {step, exp_avg, exp_avg_sq}, one param group with the hyper-parameters; plus
        `ema` = {num_updates, decay} (the EMA shadow itself is checkpointed as `model_ema.*` by ema.LitEma.bind)."""
        state = {}
        for i, (p, off) in enumerate(zip(self.params, self.buckets.offsets)):
            n = p.numel()
            state[i] = {"step": torch.tensor(float(self.step_count)),
                        "exp_avg": self.exp_avg[off:off + n].view_as(p).clone(),
                        "exp_avg_sq": self.exp_avg_sq[off:off + n].view_as(p).clone()}
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.weight_decay,
                 "amsgrad": False, "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group],
                "ema": {"num_updates": self.num_updates, "decay": self.ema_decay}}

    @torch.no_grad()
    def load_state_dict(self, sd: dict) -> None:
        group = sd["param_groups"][0]
        assert len(group["params"]) == len(self.params), "optimizer state does not match the parameter list"
        self.lr, self.betas = group["lr"], tuple(group["betas"])
        self.eps, self.weight_decay = group["eps"], group["weight_decay"]
        steps = set()
        for i, (p, off) in enumerate(zip(self.params, self.buckets.offsets)):
            st = sd["state"].get(i, sd["state"].get(str(i)))
            if st is None:
                continue
            n = p.numel()
            self.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
            steps.add(int(st["step"]))
        assert len(steps) <= 1, "per-parameter step counts differ: not a state this fused optimizer can resume"
        self.step_count = steps.pop() if steps else 0
        if "ema" in sd:
            self.num_updates, self.ema_decay = sd["ema"]["num_updates"], sd["ema"]["decay"]
        self.refresh()      # parameters may have been loaded just before: rebuild the bf16 copy, drop packed weights

    # ---- LitEma interface (ema.py:46-76) ----------------------------------------------------------------------
    def ema_views(self) -> Dict[int, torch.Tensor]:
        return {id(p): self.ema[off:off + p.numel()].view_as(p) for p, off in zip(self.params, self.buckets.offsets)}

    def store(self):
        self._stored = self.flat_param.clone()

    def copy_to(self):
        """Load the EMA weights into the model (validation / checkpointing with `ema_scope`, ddpm.py:173-186)."""
        self.flat_param.copy_(self.ema)
        self.refresh()

    def restore(self):
        self.flat_param.copy_(self._stored)
        self.refresh()
        del self._stored
