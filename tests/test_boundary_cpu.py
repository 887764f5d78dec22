"""CPU tests of the drop-in boundary against the REFERENCE's own classes and of the host logic
that mirrors the reference trainer: EMA checkpoint layout / arithmetic, `ema_scope`, the negative-conditioning branch
of `LatentDiffusion.forward`, `get_input` / `shared_step`.

The reference-hosted tests import the unmodified reference from /root/reference (with the sys.modules shims of
oracle/gen_golden.py) and are skipped where that tree does not exist (the GPU box)."""
import os
import sys

import pytest
import torch
import torch.nn as nn
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("EALDM_REFERENCE", "/root/reference")
needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "ldm")), reason="reference tree not present")

from ealdm_b200 import configs as CFG  # noqa: E402
from ealdm_b200.ddpm import LatentDiffusion  # noqa: E402
from ealdm_b200.ema import LitEma  # noqa: E402
from oracle import unet as OU  # noqa: E402


def _ref_modules():
    from oracle.gen_golden import install_shims
    install_shims()
    import ldm.models.diffusion.ddpm as ref_ddpm
    import ldm.util as ref_util
    return ref_util, ref_ddpm


# ---- the reference hosts the B200 modules -------------------------------------------------------------------------
@needs_ref
def test_reference_latent_diffusion_hosts_b200_unet_and_strict_loads_reference_state_dict():
    """The swap INTEGRATION.md describes, executed by the reference's own code: `ldm.util.instantiate_from_config` builds
    the reference `LatentDiffusion` from the shipped yaml with only `unet_config.target` edited; the result holds
    ealdm_b200.unet.UNetModel, whose state_dict strict-loads a state_dict produced by the reference UNetModel."""
    ref_util, ref_ddpm = _ref_modules()
    cfg = yaml.safe_load(open(os.path.join(REF, "configs/latent-diffusion/stdiff_cin-ldm-vq-f8.yaml")))["model"]
    assert cfg["target"] == "ldm.models.diffusion.ddpm.LatentDiffusion"
    p = cfg["params"]
    ref_unet_target = p["unet_config"]["target"]
    p["unet_config"]["target"] = "ealdm_b200.unet.UNetModel"                     # THE edit
    p["first_stage_config"] = {"target": "ldm.models.autoencoder.IdentityFirstStage"}   # (taming / lightning absent here)
    p["cond_stage_config"] = {"target": "torch.nn.Identity"}
    p["cond_stage_trainable"] = False
    p.pop("cond_stage_key", None)
    p["use_ema"] = True                                                              # the reference default
    model = ref_util.instantiate_from_config(cfg)
    assert type(model) is ref_ddpm.LatentDiffusion
    unet = model.model.diffusion_model
    assert type(unet).__module__.startswith("ealdm_b200") or "environment-aware" in type(unet).__module__
    assert unet.in_channels == 4 and unet.image_size == 32
    # a state_dict made by the REFERENCE UNetModel loads strictly, names, shapes and order included
    ref_unet = ref_util.instantiate_from_config({"target": ref_unet_target, "params": p["unet_config"]["params"]})
    sd = ref_unet.state_dict()
    assert list(sd.keys()) == list(unet.state_dict().keys())
    missing, unexpected = unet.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    # the reference's LitEma sees the same 626 trainable tensors, so `model_ema.*` checkpoints line up
    assert len(list(model.model_ema.buffers())) == 626 + 2
    # the reference's own conditioning route reaches the module (CPU: it must refuse loudly, not fall back)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model.apply_model(torch.zeros(1, 4, 32, 32), torch.zeros(1, dtype=torch.long), torch.zeros(1, 4, 512))


@needs_ref
def test_b200_first_stage_posterior_passes_the_reference_isinstance_check():
    """ddpm.py:550-557 raises NotImplementedError unless `encode()` returns the REFERENCE's DiagonalGaussianDistribution
    (or a tensor): with the reference loaded, AutoencoderKL.encode's result derives from that class."""
    ref_util, ref_ddpm = _ref_modules()
    from ldm.modules.distributions.distributions import DiagonalGaussianDistribution as RefDGD
    from ealdm_b200.autoencoder import DiagonalGaussianDistribution, posterior_class
    cls = posterior_class()
    assert issubclass(cls, RefDGD) and issubclass(cls, DiagonalGaussianDistribution)
    post = cls(torch.randn(2, 8, 4, 4, generator=torch.Generator().manual_seed(0)))
    z = ref_ddpm.LatentDiffusion.get_first_stage_encoding(type("M", (), {"scale_factor": 0.5})(), post)
    assert z.shape == (2, 4, 4, 4)
    ref = RefDGD(post.parameters)
    assert torch.equal(post.mode(), ref.mode()) and torch.equal(post.std, ref.std)
    assert torch.equal(post.kl(), ref.kl()) and torch.allclose(post.nll(post.mean), ref.nll(ref.mean))


# ---- EMA ---------------------------------------------------------------------------------------------------------------
def _toy():
    torch.manual_seed(0)
    m = nn.Sequential(nn.Linear(5, 7), nn.ReLU(), nn.Linear(7, 3))
    m[2].bias.requires_grad_(False)      # LitEma skips frozen parameters
    return m


@needs_ref
def test_lit_ema_matches_reference_layout_and_arithmetic_bit_for_bit():
    _ref_modules()
    from ldm.modules.ema import LitEma as RefEma
    a, b = _toy(), _toy()
    ea, eb = LitEma(a), RefEma(b)
    assert list(ea.state_dict().keys()) == list(eb.state_dict().keys())
    g = torch.Generator().manual_seed(1)
    for _ in range(12):                 # crosses the (1 + n) / (10 + n) warm-up of the decay
        with torch.no_grad():
            for pa, pb in zip(a.parameters(), b.parameters()):
                d = torch.randn(pa.shape, generator=g) * 0.1
                pa.add_(d)
                pb.add_(d)
        ea(a)
        eb(b)
    sa, sb = ea.state_dict(), eb.state_dict()
    for k in sb:
        assert torch.equal(sa[k], sb[k]), k
    assert int(sa["num_updates"]) == 12
    # a reference checkpoint of the shadows loads into ours, and copy_to / restore round-trip
    ea2 = LitEma(_toy())
    ea2.load_state_dict(sb, strict=True)
    assert ea2._num_updates_host == 12
    before = [p.detach().clone() for p in a.parameters()]
    ea.store(a.parameters())
    ea.copy_to(a)
    eb.copy_to(b)
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert torch.equal(pa, pb)
    ea.restore(a.parameters(), a)
    for p, q in zip(a.parameters(), before):
        assert torch.equal(p, q)


def test_latent_diffusion_use_ema_state_dict_and_ema_scope_invalidate_packed_weights():
    ld = LatentDiffusion(unet_config={"target": "ealdm_b200.unet.UNetModel", "params": dict(CFG.UNET_UNCOND)},
                         use_ema=True, **CFG.DIFFUSION)
    sd = ld.state_dict()
    ema_keys = [k for k in sd if k.startswith("model_ema.")]
    assert len(ema_keys) == 306 + 2 and "model_ema.decay" in sd and "model_ema.num_updates" in sd
    assert "model_ema.diffusion_modelinput_blocks00weight" in sd        # dots stripped, as ema.py:19-21
    unet = ld.model.diffusion_model
    unet._engine = object()            # stands for packed weights + graphs built by an earlier eval forward
    with torch.no_grad():
        ld.model_ema.diffusion_modelout2bias.fill_(3.0)
    w0 = unet.out[2].bias.detach().clone()
    with ld.ema_scope():
        assert unet._engine is None                                     # dropped on the switch to EMA weights
        assert float(unet.out[2].bias[0]) == 3.0
        unet._engine = object()
    assert unet._engine is None and torch.equal(unet.out[2].bias, w0)  # and again on the way back
    # checkpoint round trip keeps the shadows (the reference's layout: model_ema.* next to model.*)
    ld2 = LatentDiffusion(unet_config={"target": "ealdm_b200.unet.UNetModel", "params": dict(CFG.UNET_UNCOND)},
                          use_ema=True, **CFG.DIFFUSION)
    missing, unexpected = ld2.load_state_dict(ld.state_dict(), strict=True)
    assert not missing and not unexpected
    assert float(ld2.model_ema.diffusion_modelout2bias[0]) == 3.0


def test_version_fingerprint_detects_in_place_optimizer_updates():
    from ealdm_b200.unet import UNetModel
    m = UNetModel(**CFG.UNET_UNCOND)
    m._plist = list(m.parameters())
    fp = m._version_fingerprint()
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3)
    for p in m.parameters():
        p.grad = torch.zeros_like(p)
    opt.step()
    assert m._version_fingerprint() != fp


# ---- LatentDiffusion.forward: negative conditioning (ddpm.py:878-900) ---------------------------------------------------
class _Cond(nn.Module):
    """Stands for UnetCond: context = mean of the frames; the negative branch sees mixed[-1] is None."""

    def __init__(self):
        super().__init__()
        self.w = nn.Parameter(torch.ones(()))
        self.calls = []

    def forward(self, mixed):
        self.calls.append(mixed[-1] is None)
        return mixed[0].mean(dim=(1, 2, 3)).reshape(-1, 1, 1).expand(-1, 4, 512) * self.w


def test_forward_builds_negative_conditioning_like_the_reference():
    ld = LatentDiffusion(unet_config={"target": "ealdm_b200.unet.UNetModel", "params": dict(CFG.UNET_STDIFF)},
                         cond_stage_config={"target": "torch.nn.Identity"}, cond_stage_trainable=True,
                         cond_stage_key="mixed", conditioning_key="crossattn", **CFG.DIFFUSION)
    ld.cond_stage_model = _Cond()
    seen = {}

    def fake_p_losses(x, c, t, *a, **k):
        seen["c"], seen["t"] = c, t
        return torch.zeros(()), {}

    ld.p_losses = fake_p_losses
    frames, neg = torch.full((2, 3, 8, 8), 1.0), torch.full((2, 3, 8, 8), -2.0)
    mixed = [frames, torch.zeros(2, 1, 1), torch.zeros(2, 1, 16), torch.zeros(2, 1), neg]
    ld(torch.zeros(2, 4, 32, 32), mixed)
    c = seen["c"]
    assert c.shape == (4, 4, 512)
    assert torch.all(c[:2] == -2.0) and torch.all(c[2:] == 1.0)       # cat([c_neg, c]): negatives first
    assert ld.cond_stage_model.calls == [True, False]                  # c_neg[-1] is None, then the positive branch
    assert mixed[0] is frames and mixed[-1] is neg                     # the caller's batch is not mutated
    assert seen["t"].shape == (2,) and seen["t"].dtype == torch.long
    # configure_optimizers: UNet + conditioner parameters (ddpm.py:1409-1415)
    ld.learning_rate = 1e-4
    opt = ld.configure_optimizers()
    n_params = sum(len(g["params"]) for g in opt.param_groups)
    assert n_params == 626 + 1
