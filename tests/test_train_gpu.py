"""Training step (BASELINE.json configs[4]): LatentDiffusion.p_losses forward +
the hand-written backward, against golden gradients made by the UNMODIFIED reference (autograd through
its own modules on CPU fp32; oracle/gen_golden_grads.py -> tests/golden/p_losses_grads.pt).

Tolerances (stated here; the north star only fixes the eps tolerances 1e-4 / 1e-2):
  fp32 mode: every parameter gradient within 5e-4 relative L2 (norm, random projection, and full tensors);
  bf16 mode: within 5e-2 relative L2 (bf16 GEMM operands in both passes, fp32 accumulation, fp32 residual
  and gradient streams) -- the same order as torch.autocast(bf16) training.
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from ealdm_b200 import configs as CFG  # noqa: E402
from ealdm_b200.ddpm import LatentDiffusion  # noqa: E402
from oracle import unet as OU  # noqa: E402  (checker only)

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = {"fp32": 5e-4, "bf16": 5e-2}


def gold(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def direction(i, shape):
    return torch.randn(shape, generator=torch.Generator().manual_seed(1000 + i))


_ld = {}


def make_ld():
    if "ld" not in _ld:
        ld = LatentDiffusion(unet_config={"target": "ealdm_b200.unet.UNetModel", "params": dict(CFG.UNET_STDIFF)},
                             cond_stage_config={"target": "torch.nn.Identity"}, conditioning_key="crossattn",
                             **CFG.DIFFUSION)
        sd = OU.synthetic_state_dict(OU.unet_param_shapes(CFG.UNET_STDIFF), seed=2)
        ld.model.diffusion_model.load_state_dict(sd, strict=True)
        _ld["ld"] = ld.cuda()
    return _ld["ld"]


def run_step(mode):
    G = gold("p_losses.pt")
    ld = make_ld().train()
    unet = ld.model.diffusion_model.set_compute_dtype(mode)
    for p in unet.parameters():
        p.grad = None
    c2 = G["cond2"].cuda().requires_grad_(True)
    loss, _ = ld.p_losses(G["x0"].cuda(), c2, G["t"].cuda(), noise=G["noise"].cuda())
    loss.backward()
    return ld, unet, loss, c2


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_p_losses_gradients_vs_reference_golden(mode):
    GG = gold("p_losses_grads.pt")
    ld, unet, loss, c2 = run_step(mode)
    tol = TOL[mode]
    assert abs(float(loss) - float(GG["loss"])) <= tol * abs(float(GG["loss"]))
    names = [n for n, _ in unet.named_parameters()]
    assert names == GG["names"]
    worst_n, worst_p = (0.0, ""), (0.0, "")
    for i, (name, p) in enumerate(unet.named_parameters()):
        assert p.grad is not None, f"no gradient for {name}"
        gr = p.grad.double().cpu()
        ref_n = GG["norm"][name]
        en = abs(float(gr.norm()) - ref_n) / max(ref_n, 1e-30)
        # <g - g_ref, r> ~ N(0, ||g - g_ref||^2) for a unit-variance direction r: 4 sigma
        ep = abs(float((gr * direction(i, gr.shape).double()).sum()) - GG["proj"][name]) / max(ref_n, 1e-30) / 4.0
        worst_n = max(worst_n, (en, name))
        worst_p = max(worst_p, (ep, name))
    print(f"train {mode}: worst |norm| error {worst_n[0]:.3e} ({worst_n[1]}), worst projection error "
          f"{worst_p[0]:.3e} ({worst_p[1]})")
    assert worst_n[0] < tol, worst_n
    assert worst_p[0] < tol, worst_p
    sd = dict(unet.named_parameters())
    worst_f = (0.0, "")
    for name, ref in GG["full"].items():
        worst_f = max(worst_f, (rel_l2(sd[name].grad, ref), name))
    print(f"train {mode}: worst full-tensor gradient rel_l2 {worst_f[0]:.3e} ({worst_f[1]})")
    assert worst_f[0] < tol, worst_f
    e = rel_l2(c2.grad, GG["dcond"])
    print(f"train {mode}: d(loss)/d(conditioning) rel_l2 {e:.3e}")
    assert e < tol
    ld.eval()


def test_backward_is_deterministic_and_accumulates():
    ld, unet, loss, _ = run_step("bf16")
    g1 = {n: p.grad.clone() for n, p in unet.named_parameters()}
    ld, unet, loss2, _ = run_step("bf16")
    assert float(loss) == float(loss2)
    for n, p in unet.named_parameters():
        assert torch.equal(p.grad, g1[n]), n
    # a second backward without zeroing accumulates (PyTorch .grad semantics)
    G = gold("p_losses.pt")
    l3, _ = ld.p_losses(G["x0"].cuda(), G["cond2"].cuda(), G["t"].cuda(), noise=G["noise"].cuda())
    l3.backward()
    for n, p in list(unet.named_parameters())[::37]:
        assert rel_l2(p.grad, 2 * g1[n]) < 1e-6, n
    ld.eval()


def test_adamw_steps_reduce_the_loss():
    """Three AdamW steps (the reference's optimizer, ddpm.py:1409-1431) on one fixed batch."""
    G = gold("p_losses.pt")
    ld = make_ld().train()
    unet = ld.model.diffusion_model.set_compute_dtype("bf16")
    saved = {n: p.detach().clone() for n, p in unet.named_parameters()}
    opt = torch.optim.AdamW(unet.parameters(), lr=2e-5)
    x0, c2, t, noise = G["x0"].cuda(), G["cond2"].cuda(), G["t"].cuda(), G["noise"].cuda()
    losses = []
    for _ in range(4):
        opt.zero_grad(set_to_none=True)
        loss, _ = ld.p_losses(x0, c2, t, noise=noise)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    print("losses:", losses)
    assert losses[-1] < losses[0]
    with torch.no_grad():
        for n, p in unet.named_parameters():
            p.copy_(saved[n])
    unet.invalidate_packed()
    ld.eval()


@pytest.mark.parametrize("use_graph", [False, True, "segments"])
def test_fused_step_matches_autograd_path(use_graph):
    """FusedTrainStep (the CUDA-graph-capturable trainer loop) == p_losses + loss.backward(), bit for bit."""
    from ealdm_b200.train import FusedTrainStep
    ld, unet, loss, _ = run_step("bf16")
    ref = {n: p.grad.clone() for n, p in unet.named_parameters()}
    G = gold("p_losses.pt")
    fused = FusedTrainStep(ld, use_graph=bool(use_graph))
    fused.force_segments = use_graph == "segments"     # the multi-GPU capture layout (5 graphs), on one rank
    args = (G["x0"].cuda(), G["cond2"].cuda(), G["t"].cuda(), G["noise"].cuda())
    for it in range(3 if use_graph else 1):      # graph: capture, then two replays
        fused.buckets.zero_()
        l2 = fused(*args)
        fused.buckets.finish()
        assert float(l2) == float(loss)
        for n, p in unet.named_parameters():
            assert torch.equal(p.grad, ref[n]), (it, n)
    if use_graph == "segments":
        assert len(next(iter(fused._graphs.values()))[0]) == 5
    unet.grad_ready_hook = None
    for p in unet.parameters():
        p.grad = None
    ld.eval()


def test_fused_adamw_ema_matches_torch_adamw_and_litema():
    """ealdm_adamw_ema_step == torch.optim.AdamW (single-tensor reference math) + the LitEma update, on the flat
    buffers of a small UNet; the bf16 copy equals the cast of the new weights."""
    from ealdm_b200.optim import FusedAdamWEMA
    from ealdm_b200.parallel import GradBuckets
    from ealdm_b200.unet import UNetModel
    tiny = dict(image_size=8, in_channels=4, model_channels=64, out_channels=4, num_res_blocks=1,
                attention_resolutions=[1, 2], channel_mult=(1, 2), num_head_channels=32,
                use_spatial_transformer=True, transformer_depth=1, context_dim=64)
    torch.manual_seed(0)
    unet = UNetModel(**tiny).cuda()
    ref = UNetModel(**tiny).cuda()
    ref.load_state_dict(unet.state_dict())
    gb = GradBuckets(unet, bucket_mb=1.0)
    opt = FusedAdamWEMA(gb, lr=3e-4, weight_decay=1e-2, ema_decay=0.9999)
    ropt = torch.optim.AdamW(ref.parameters(), lr=3e-4, weight_decay=1e-2, foreach=False, fused=False)
    shadow = {n: p.detach().clone() for n, p in ref.named_parameters()}
    gen = torch.Generator(device="cuda").manual_seed(5)
    for step in range(1, 4):
        gb.zero_()
        for (n, p), (_, q) in zip(unet.named_parameters(), ref.named_parameters()):
            gr = torch.randn(p.shape, generator=gen, device="cuda") * 0.1
            p.grad.copy_(gr)
            q.grad = gr.clone()
        gb.finish()
        opt.step()
        ropt.step()
        decay = min(0.9999, (1 + step) / (10 + step))
        for n, q in ref.named_parameters():
            shadow[n].sub_((1.0 - decay) * (shadow[n] - q.detach()))
    ema = opt.ema_views()
    worst = 0.0
    for (n, p), (_, q) in zip(unet.named_parameters(), ref.named_parameters()):
        worst = max(worst, rel_l2(p.detach(), q.detach()), rel_l2(ema[id(p)], shadow[n]))
        assert torch.equal(p._bf16_view, p.detach().to(torch.bfloat16)), n
    print(f"fused AdamW+EMA vs torch: worst rel_l2 {worst:.3e}")
    assert worst < 1e-6


def test_eval_forward_after_weight_updates_uses_the_new_weights():
    """the packed inference engine (and its CUDA graphs) is a cache of the parameters.  After an optimizer
    step (version counters), a write through `.data` + revalidate_packed() (the reference LitEma.copy_to path) and
    FusedAdamWEMA.step(), an eval forward must equal the forward of a freshly packed engine."""
    from ealdm_b200.optim import FusedAdamWEMA
    from ealdm_b200.parallel import GradBuckets
    from ealdm_b200.unet import UNetModel
    tiny = dict(image_size=8, in_channels=4, model_channels=128, out_channels=4, num_res_blocks=1,
                attention_resolutions=[1, 2], channel_mult=(1, 2), num_head_channels=32,
                use_spatial_transformer=True, transformer_depth=1, context_dim=64)
    torch.manual_seed(0)
    unet = UNetModel(**tiny).cuda().eval().enable_cuda_graph(True)
    with torch.no_grad():
        for p in unet.parameters():      # the zero-initialised output convs would hide every upstream change
            if p.dim() >= 2 and float(p.abs().max()) == 0.0:
                p.normal_(std=0.05)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 4, 8, 8, generator=g).cuda()
    c = torch.randn(2, 4, 64, generator=g).cuda()
    t = torch.tensor([5, 900]).cuda()

    def fresh():
        m = UNetModel(**tiny).cuda().eval()
        m.load_state_dict(unet.state_dict())
        return m(x, t, context=c)

    y0 = unet(x, t, context=c)
    y0 = unet(x, t, context=c)           # graph replay
    assert torch.equal(y0, fresh())
    # (1) a torch optimizer step: caught by the version counters on the next forward
    opt = torch.optim.SGD(unet.parameters(), lr=0.05)
    for p in unet.parameters():
        p.grad = torch.ones_like(p)
    opt.step()
    y1 = unet(x, t, context=c)
    assert not torch.equal(y1, y0) and torch.equal(y1, fresh())
    # (2) a write through .data (invisible to version counters): caught by the samplers' revalidate_packed()
    with torch.no_grad():
        for p in unet.parameters():
            p.data.mul_(1.01)
    assert unet.revalidate_packed() is False
    y2 = unet(x, t, context=c)
    assert not torch.equal(y2, y1) and torch.equal(y2, fresh())
    assert unet.revalidate_packed() is True
    # (3) the fused optimizer invalidates explicitly
    for p in unet.parameters():
        p.grad = None
    gb = GradBuckets(unet, bucket_mb=1.0)
    fopt = FusedAdamWEMA(gb, lr=1e-2)
    y3 = unet(x, t, context=c)
    gb.flat.fill_(0.5)
    fopt.step()
    assert unet._engine is None
    y4 = unet(x, t, context=c)
    assert not torch.equal(y4, y3) and torch.equal(y4, fresh())
    # (4) EMA weights: copy_to / restore drop the packed weights as well
    fopt.store()
    fopt.copy_to()
    y5 = unet(x, t, context=c)
    assert torch.equal(y5, fresh()) and not torch.equal(y5, y4)
    fopt.restore()
    assert torch.equal(unet(x, t, context=c), y4)


def test_fused_optimizer_and_ema_checkpoint_round_trip():
    """FusedAdamWEMA.state_dict() has torch.optim.AdamW's per-parameter layout; LitEma.bind() exposes the fused EMA
    buffer as `model_ema.*` (the reference's checkpoint keys); a resumed run continues bit-identically."""
    from ealdm_b200.ema import LitEma
    from ealdm_b200.optim import FusedAdamWEMA
    from ealdm_b200.parallel import GradBuckets
    from ealdm_b200.unet import UNetModel
    tiny = dict(image_size=8, in_channels=4, model_channels=128, out_channels=4, num_res_blocks=1,
                attention_resolutions=[1, 2], channel_mult=(1, 2), num_head_channels=32,
                use_spatial_transformer=True, transformer_depth=1, context_dim=64)

    def make(seed):
        torch.manual_seed(seed)
        unet = UNetModel(**tiny).cuda()
        gb = GradBuckets(unet, bucket_mb=1.0)
        opt = FusedAdamWEMA(gb, lr=3e-4)
        ema = LitEma(unet).cuda().bind(opt, unet)
        return unet, gb, opt, ema

    def steps(gb, opt, n, seed):
        gen = torch.Generator(device="cuda").manual_seed(seed)
        for _ in range(n):
            gb.flat.copy_(torch.randn(gb.flat.shape, generator=gen, device="cuda") * 0.1)
            opt.step()

    unet, gb, opt, ema = make(0)
    steps(gb, opt, 3, 11)
    ck = {"model": {k: v.clone() for k, v in unet.state_dict().items()}, "ema": ema.state_dict(), "opt": opt.state_dict()}
    ck["ema"] = {k: v.clone() for k, v in ck["ema"].items()}
    ref_sd = torch.optim.AdamW(unet.parameters()).state_dict()
    assert set(ck["opt"]["param_groups"][0]) >= {"lr", "betas", "eps", "weight_decay", "params"}
    assert ck["opt"]["param_groups"][0]["params"] == ref_sd["param_groups"][0]["params"]
    assert set(ck["opt"]["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
    assert int(ck["ema"]["num_updates"]) == 3 and len(ck["ema"]) == len(list(unet.parameters())) + 2
    steps(gb, opt, 2, 12)                 # the continued run
    want = {k: v.clone() for k, v in unet.state_dict().items()}
    want_ema = {n: opt.ema_views()[id(p)].clone() for n, p in unet.named_parameters()}
    # resume from the checkpoint in a new process' worth of objects
    unet2, gb2, opt2, ema2 = make(1)
    unet2.load_state_dict(ck["model"])
    ema2.load_state_dict(ck["ema"])
    opt2.load_state_dict(ck["opt"])
    assert opt2.step_count == 3 and opt2.num_updates == 3
    steps(gb2, opt2, 2, 12)
    for k, v in unet2.state_dict().items():
        assert torch.equal(v, want[k]), k
    views2 = opt2.ema_views()      # (the flat buffers also hold alignment padding, which no checkpoint carries)
    for n, p in unet2.named_parameters():
        assert torch.equal(views2[id(p)], want_ema[n]), n
