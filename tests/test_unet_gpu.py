"""End-to-end parity of the CUDA path (through the module API -> C ABI) against golden vectors made
by the reference's own modules (tests/golden, see oracle/gen_golden.py) and against the CPU oracle.

Tolerances are the north star's: per-step eps / x_prev within 1e-4 relative L2 in fp32 mode and
1e-2 in bf16 mode; schedule bit-exact (tests/test_host_cpu.py)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from ealdm_b200 import configs as CFG  # noqa: E402
from ealdm_b200.ddim import DDIMSampler  # noqa: E402
from ealdm_b200.ddpm import LatentDiffusion  # noqa: E402
from ealdm_b200.unet import UNetModel  # noqa: E402
from oracle import unet as OU  # noqa: E402  (checker only)

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = {"fp32": 1e-4, "bf16": 1e-2}


def gold(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


_cache = {}


def make_ld(kind):
    """LatentDiffusion with the B200 UNet and the synthetic weights the golden files were made with."""
    if kind in _cache:
        return _cache[kind]
    ucfg = CFG.UNET_STDIFF if kind == "stdiff" else CFG.UNET_UNCOND
    target = "ealdm_b200.unet.UNetModel"
    ld = LatentDiffusion(unet_config={"target": target, "params": dict(ucfg)},
                         cond_stage_config={"target": "torch.nn.Identity"} if kind == "stdiff" else "__is_unconditional__",
                         conditioning_key="crossattn" if kind == "stdiff" else None, **CFG.DIFFUSION)
    sd = OU.synthetic_state_dict(OU.unet_param_shapes(ucfg), seed=2 if kind == "stdiff" else 1)
    ld.model.diffusion_model.load_state_dict(sd, strict=True)
    ld = ld.cuda().eval()
    _cache[kind] = ld
    return ld


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("kind", ["uncond", "stdiff"])
def test_unet_forward_vs_reference_golden(kind, mode):
    G = gold(f"unet_{kind}_fwd.pt")
    ld = make_ld(kind)
    unet = ld.model.diffusion_model.set_compute_dtype(mode)
    ctx = None if G["context"] is None else G["context"].cuda()
    eps = unet(G["x"].cuda(), G["t"].cuda(), context=ctx)
    assert eps.shape == G["eps"].shape and eps.dtype == torch.float32
    err = rel_l2(eps, G["eps"])
    print(f"unet {kind} {mode}: rel_l2 = {err:.3e}")
    assert err < TOL[mode]


@pytest.mark.parametrize("kind", ["uncond", "stdiff"])
def test_unet_forward_bench_schedules_vs_reference_golden(kind):
    """The schedules the BENCHMARK batch selects (256-column N tiles, CTA pairs, wide epilogue passes) forced onto the
    golden batch of 2, where the wave-count rule alone would pick 128-column tiles: same tolerance, and bit-identical to
    the default schedule (every schedule accumulates the same k-blocks in the same order)."""
    from ealdm_b200 import _lib as L
    G = gold(f"unet_{kind}_fwd.pt")
    unet = make_ld(kind).model.diffusion_model.set_compute_dtype("bf16")
    ctx = None if G["context"] is None else G["context"].cuda()
    base = unet(G["x"].cuda(), G["t"].cuda(), context=ctx)
    lib = L.load()
    prev = [lib.ealdm_tc_set_option(o, v) for o, v in ((L.TC_OPT_BN, 256), (L.TC_OPT_CTA2, 2))]
    try:
        eps = unet(G["x"].cuda(), G["t"].cuda(), context=ctx)
        torch.cuda.synchronize()
    finally:
        lib.ealdm_tc_set_option(L.TC_OPT_BN, prev[0])
        lib.ealdm_tc_set_option(L.TC_OPT_CTA2, prev[1])
    err = rel_l2(eps, G["eps"])
    print(f"unet {kind} bf16, forced BN=256 + CTA pairs: rel_l2 = {err:.3e}")
    assert err < TOL["bf16"]
    assert torch.equal(eps, base)


def test_unet_forward_with_folded_layer_norm(monkeypatch):
    """The opt-in LayerNorm fold (EALDM_LN_FOLD: producer row statistics + consumer epilogue correction instead of
    LayerNorm passes): same north-star tolerance, and the same bits under every conv schedule (the producer's partial
    sums are per 32 columns whatever the N tile)."""
    import ealdm_b200.unet as U
    from ealdm_b200 import _lib as L
    monkeypatch.setattr(U.UNetEngine, "_fold_ln", True)
    G = gold("unet_stdiff_fwd.pt")
    unet = make_ld("stdiff").model.diffusion_model.set_compute_dtype("bf16")
    x, t, c = G["x"].cuda(), G["t"].cuda(), G["context"].cuda()
    base = unet(x, t, context=c)
    err = rel_l2(base, G["eps"])
    print(f"unet stdiff bf16, LayerNorm folded: rel_l2 = {err:.3e}")
    assert err < TOL["bf16"]
    lib = L.load()
    prev = [lib.ealdm_tc_set_option(o, v) for o, v in ((L.TC_OPT_BN, 256), (L.TC_OPT_CTA2, 2))]
    try:
        assert torch.equal(unet(x, t, context=c), base)
    finally:
        lib.ealdm_tc_set_option(L.TC_OPT_BN, prev[0])
        lib.ealdm_tc_set_option(L.TC_OPT_CTA2, prev[1])


def test_unet_forward_programmatic_dependent_launch_is_bit_identical():
    """ealdm_set_pdl: the kernels of the forward are launched with programmatic stream serialization (each one waits on
    griddepcontrol.wait before its first global access).  Eager and CUDA-graph replays with the attribute must equal
    the plainly serialised launches bit for bit -- a kernel that starts before its predecessor's writes are visible
    would show up here."""
    from ealdm_b200 import _lib as L
    G = gold("unet_stdiff_fwd.pt")
    unet = make_ld("stdiff").model.diffusion_model.set_compute_dtype("bf16")
    x, t, c = G["x"].cuda(), G["t"].cuda(), G["context"].cuda()
    x = torch.cat([x] * 8); t = torch.cat([t] * 8); c = torch.cat([c] * 8)     # 16 samples: several tiles per SM
    lib = L.load()
    prev = lib.ealdm_set_pdl(0)
    try:
        base = unet(x, t, context=c)
        lib.ealdm_set_pdl(1)
        for _ in range(3):
            assert torch.equal(unet(x, t, context=c), base)
        unet.enable_cuda_graph(True)
        for _ in range(3):
            assert torch.equal(unet(x, t, context=c), base)
        unet.enable_cuda_graph(False)
    finally:
        lib.ealdm_set_pdl(prev)
    assert rel_l2(base[:2], G["eps"]) < TOL["bf16"]


def test_unet_forward_guidance_pair_shares_the_prefix_bit_identical():
    """Classifier-free guidance evaluates cat([x] * 2), cat([t] * 2), cat([uc, c]) (ddim.py:176-179): under
    `unet.cfg_pair()` the layers in front of the first cross-attention run once on the first half of the batch and the
    token stream is duplicated.  Must equal the plain forward of the duplicated batch bit for bit, eager and graphed,
    and (with `sampling_scope`) a context that is the same tensor is projected once, a modified one again."""
    from ealdm_b200 import ops
    G = gold("unet_stdiff_fwd.pt")
    unet = make_ld("stdiff").model.diffusion_model.set_compute_dtype("bf16")
    g = torch.Generator().manual_seed(5)
    x = torch.cat([G["x"].cuda()] * 4)                   # 8 images
    x = x + 0.1 * torch.randn(x.shape, generator=g).cuda()
    t = torch.cat([G["t"].cuda()] * 4)
    c = torch.randn(16, 4, 512, generator=g).cuda()      # uc rows, then c rows
    x2, t2 = torch.cat([x] * 2), torch.cat([t] * 2)
    base = unet(x2, t2, context=c)
    assert not torch.equal(base[:8], base[8:])           # the halves do differ (through the context only)
    with unet.cfg_pair():
        n0 = ops.launch_count()
        shared = unet(x2, t2, context=c)
        n_shared = ops.launch_count() - n0
    n0 = ops.launch_count()
    again = unet(x2, t2, context=c)
    n_plain = ops.launch_count() - n0
    assert torch.equal(shared, base) and torch.equal(again, base)
    assert n_shared == n_plain                            # same launches, the prefix on half the rows
    for graph in (False, True):
        unet.enable_cuda_graph(graph)
        with unet.sampling_scope():
            with unet.cfg_pair():
                a = unet(x2, t2, context=c)
                n0 = ops.launch_count()
                b = unet(x2, t2, context=c)               # same context tensor: not projected again
                n_reuse = ops.launch_count() - n0
                c.mul_(1.0)                               # in-place write: version bump, same values
                n0 = ops.launch_count()
                d = unet(x2, t2, context=c)
                n_again = ops.launch_count() - n0
                c2 = torch.cat([c[8:], c[:8]])            # another tensor: uc and c rows swapped
                e = unet(x2, t2, context=c2)
        assert torch.equal(a, base) and torch.equal(b, base) and torch.equal(d, base)
        assert torch.equal(e, torch.cat([base[8:], base[:8]]))
        if not graph:
            assert n_again - n_reuse >= 2                 # the context copy + projection GEMMs came back
        outside = unet(x2, t2, context=c2)                # outside the scope nothing is remembered
        assert torch.equal(outside, e)
    unet.enable_cuda_graph(False)


def test_guided_ddim_sampling_is_bit_identical_with_and_without_the_sampler_hints(monkeypatch):
    """End to end: `DDIMSampler.sample` with classifier-free guidance through the hinted path (shared prefix of the
    cond / uncond halves, context projected once per loop, CUDA-graph replay) against the same call with the hints
    switched off -- the latents must be the same bits (the hints remove duplicated work, they change no arithmetic)."""
    ld = make_ld("stdiff")
    unet = ld.model.diffusion_model.set_compute_dtype("bf16")
    g = torch.Generator().manual_seed(11)
    B = 6
    x_T = torch.randn(B, 4, 32, 32, generator=g).cuda()
    c = torch.randn(B, 4, 512, generator=g).cuda()
    uc = torch.randn(B, 4, 512, generator=g).cuda()

    def run():
        torch.manual_seed(3)      # the per-step noise draw (eta = 0.5)
        z, _ = DDIMSampler(ld).sample(S=6, batch_size=B, shape=(4, 32, 32), conditioning=c, eta=0.5, x_T=x_T,
                                      verbose=False, unconditional_guidance_scale=2.0, unconditional_conditioning=uc)
        return z

    outs = {}
    for graph in (False, True):
        unet.enable_cuda_graph(graph)
        outs["hints", graph] = run()
        outs["hints again", graph] = run()
        monkeypatch.setenv("EALDM_NO_CFG_SHARE", "1")
        monkeypatch.setenv("EALDM_NO_CTX_REUSE", "1")
        outs["plain", graph] = run()
        monkeypatch.delenv("EALDM_NO_CFG_SHARE")
        monkeypatch.delenv("EALDM_NO_CTX_REUSE")
    unet.enable_cuda_graph(False)
    base = outs["plain", False]
    assert torch.isfinite(base).all()
    for k, v in outs.items():
        assert torch.equal(v, base), k


def test_unet_is_deterministic_and_batch_independent():
    G = gold("unet_stdiff_fwd.pt")
    unet = make_ld("stdiff").model.diffusion_model.set_compute_dtype("bf16")
    x, t, c = G["x"].cuda(), G["t"].cuda(), G["context"].cuda()
    a = unet(x, t, context=c)
    b = unet(x, t, context=c)
    assert torch.equal(a, b)
    # sample 0 alone == sample 0 inside the batch (no cross-sample leakage through tiles / stats)
    a0 = unet(x[:1].contiguous(), t[:1].contiguous(), context=c[:1].contiguous())
    assert rel_l2(a0, a[:1]) < 2e-3
    # ragged batch (odd N exercises partial tiles at the 8x8 level)
    x3 = torch.cat([x, x[:1]]); t3 = torch.cat([t, t[:1]]); c3 = torch.cat([c, c[:1]])
    a3 = unet(x3, t3, context=c3)
    assert rel_l2(a3[:2], a) < 2e-3 and rel_l2(a3[2:], a[:1]) < 2e-3


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_ddim_config1_trajectory_vs_reference_golden(mode):
    """BASELINE.json configs[0]: uncond UNet, 10-step DDIM, batch 4, eta 0."""
    G = gold("ddim_traj.pt")["config1_uncond_B4_S10_eta0"]
    ld = make_ld("uncond")
    ld.model.diffusion_model.set_compute_dtype(mode)
    sampler = DDIMSampler(ld)
    xs, ps = [], []
    orig = sampler.p_sample_ddim

    def wrap(*a, **k):
        out = orig(*a, **k)
        xs.append(out[0]); ps.append(out[1])
        return out

    sampler.p_sample_ddim = wrap
    samples, inter = sampler.sample(S=10, batch_size=4, shape=(4, 32, 32), eta=0.0, x_T=G["x_T"].cuda(), verbose=False)
    errs = [rel_l2(xs[i], G["x_prev"][i]) for i in range(10)]
    errp = [rel_l2(ps[i], G["pred_x0"][i]) for i in range(10)]
    print(f"ddim config1 {mode}: x_prev rel_l2 per step = {['%.2e' % e for e in errs]}")
    print(f"ddim config1 {mode}: pred_x0 rel_l2 per step = {['%.2e' % e for e in errp]}")
    assert max(errs) < TOL[mode] and max(errp) < TOL[mode] * 3
    assert rel_l2(samples, G["samples"]) < TOL[mode]
    assert len(inter["x_inter"]) == 3  # x_T, index 9 (first step) and index 0 (log_every_t=100)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_ddim_stdiff_cfg_eta1_vs_reference_golden(mode):
    G = gold("ddim_traj.pt")["stdiff_B2_S10_eta1_cfg2"]
    ld = make_ld("stdiff")
    ld.model.diffusion_model.set_compute_dtype(mode)
    noises = [n.cuda() for n in G["noise"]]
    sampler = DDIMSampler(ld, noise_fn=lambda shape, device: noises.pop(0))
    xs = []
    orig = sampler.p_sample_ddim

    def wrap(*a, **k):
        out = orig(*a, **k)
        xs.append(out[0])
        return out

    sampler.p_sample_ddim = wrap
    samples, _ = sampler.sample(S=10, batch_size=2, shape=(4, 32, 32), conditioning=G["cond"].cuda(), eta=1.0,
                                x_T=G["x_T"].cuda(), verbose=False, unconditional_guidance_scale=2.0,
                                unconditional_conditioning=G["uc"].cuda())
    errs = [rel_l2(xs[i], G["x_prev"][i]) for i in range(10)]
    print(f"ddim stdiff cfg {mode}: x_prev rel_l2 per step = {['%.2e' % e for e in errs]}")
    assert max(errs) < TOL[mode]
    assert rel_l2(samples, G["samples"]) < TOL[mode]


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_p_losses_vs_reference_golden(mode):
    G = gold("p_losses.pt")
    ld = make_ld("stdiff")
    ld.model.diffusion_model.set_compute_dtype(mode)
    xq = ld.q_sample(G["x0"].cuda(), G["t"].cuda(), G["noise"].cuda())
    assert torch.equal(xq.cpu(), G["q_sample"])          # bit-exact elementwise kernel
    loss, d = ld.p_losses(G["x0"].cuda(), G["cond2"].cuda(), G["t"].cuda(), noise=G["noise"].cuda())
    tol = 2e-4 if mode == "fp32" else 2e-2
    assert abs(float(loss) - float(G["loss"])) <= tol * abs(float(G["loss"]))
    assert abs(float(d["val/loss_vlb"]) - float(G["loss_dict"]["val/loss_vlb"])) <= tol * abs(float(G["loss_dict"]["val/loss_vlb"]))


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("case", ["uncond_B2_S10", "stdiff_B2_S8_cfg2"])
def test_plms_trajectory_vs_reference_golden(case, mode):
    """PLMSSampler against the reference PLMSSampler's trajectory: pseudo improved Euler
    first step (two UNet evaluations), then Adams-Bashforth orders 2-4 on the eps history."""
    from ealdm_b200.plms import PLMSSampler
    G = gold("plms_traj.pt")[case]
    ld = make_ld("uncond" if case.startswith("uncond") else "stdiff")
    ld.model.diffusion_model.set_compute_dtype(mode)
    sampler = PLMSSampler(ld)
    xs, ps = [], []
    orig = sampler.p_sample_plms

    def wrap(*a, **k):
        out = orig(*a, **k)
        xs.append(out[0]); ps.append(out[1])
        return out

    sampler.p_sample_plms = wrap
    cuda = lambda t: None if t is None else t.cuda()  # noqa: E731
    samples, inter = sampler.sample(S=G["S"], batch_size=2, shape=(4, 32, 32), conditioning=cuda(G["cond"]), eta=0.0,
                                    x_T=G["x_T"].cuda(), verbose=False, unconditional_guidance_scale=G["ugs"],
                                    unconditional_conditioning=cuda(G["uc"]))
    errs = [rel_l2(xs[i], G["x_prev"][i]) for i in range(G["S"])]
    print(f"plms {case} {mode}: x_prev rel_l2 per step = {['%.2e' % e for e in errs]}")
    tol = TOL[mode] * (1 if mode == "fp32" else 2)
    assert max(errs) < tol and rel_l2(samples, G["samples"]) < tol
    assert max(rel_l2(ps[i], G["pred_x0"][i]) for i in range(G["S"])) < tol * 3
    with pytest.raises(ValueError):
        sampler.make_schedule(10, ddim_eta=0.5, verbose=False)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_ddpm_ancestral_sampling_vs_reference_golden(mode):
    """LatentDiffusion.p_sample_loop (ddpm.py:1190-1247): the last six ancestral steps incl. the t == 0 branch."""
    G = gold("ddpm_ancestral.pt")
    ld = make_ld("stdiff")
    ld.model.diffusion_model.set_compute_dtype(mode)
    assert bool(ld.clip_denoised) == G["clip_denoised"]
    imgs = []
    out = ld.p_sample_loop(G["cond"].cuda(), (2, 4, 32, 32), x_T=G["x_T"].cuda(), verbose=False, timesteps=6,
                           noises=[n.cuda() for n in G["noise"]], img_callback=lambda img, i: imgs.append(img))
    errs = [rel_l2(imgs[k], G["imgs"][k]) for k in range(6)]
    print(f"ddpm ancestral {mode}: x_prev rel_l2 per step = {['%.2e' % e for e in errs]}")
    assert max(errs) < TOL[mode] and rel_l2(out, G["out"]) < TOL[mode]
