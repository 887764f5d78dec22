"""Deterministic on-device 'random-init' weights for benchmarks (no checkpoints, no network).

Distribution = BASELINE.md section 4: the reference constructors' initialisation (kaiming-uniform
U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for conv / linear weights and biases, norm gains 1 and biases 0)
plus N(0, 0.02) on the tensors the reference zero-initialises (ResBlock.out_layers.3, *.proj_out,
out.2 -- openaimodel.py:229-231,312,685; attention.py:244-248).  A freshly constructed reference
UNet outputs exactly 0; with this re-randomisation eps has std ~0.6 and no kernel sees trivial data."""
from __future__ import annotations

import math

import torch


def _zero_init(wname: str) -> bool:
    return wname.endswith("out_layers.3.weight") or wname.endswith("proj_out.weight") or wname == "out.2.weight"


@torch.no_grad()
def init_synthetic_(module: torch.nn.Module, seed: int = 0) -> torch.nn.Module:
    dev = next(module.parameters()).device
    gen = torch.Generator(device=dev).manual_seed(seed)
    params = dict(module.named_parameters())
    for name, p in params.items():
        if name.endswith(".bias"):
            w = params.get(name[:-5] + ".weight")
            if w is None or w.dim() == 1 or _zero_init(name[:-5] + ".weight"):
                p.zero_()
            else:
                bound = 1.0 / math.sqrt(w[0].numel())
                p.copy_((torch.rand(p.shape, generator=gen, device=dev) * 2 - 1) * bound)
        elif p.dim() == 1:
            p.fill_(1.0)
        elif _zero_init(name):
            p.copy_(0.02 * torch.randn(p.shape, generator=gen, device=dev))
        else:
            bound = 1.0 / math.sqrt(p[0].numel())
            p.copy_((torch.rand(p.shape, generator=gen, device=dev) * 2 - 1) * bound)
    for m in module.modules():
        if hasattr(m, "invalidate_packed"):
            m.invalidate_packed()
    return module
