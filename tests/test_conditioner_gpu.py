"""The EALDM conditioner `UnetCond` on the CUDA path (ealdm_b200/conditioner.py -> C ABI) against golden vectors made by
the reference's own UnetCond + VQModelInterface (oracle/gen_golden_cond.py -> tests/golden/conditioner.pt).

Tolerances: everything downstream of the encoder is fp32 -> 1e-4 relative L2 when the encoder output is given or runs
in fp32 mode; with the first stage in bf16 mode the context inherits the encoder's bf16 error -> 3e-2 (the encoder's
own bf16 bound against the fp32 reference is 1e-2 on z, and AdaIN's instance normalisation amplifies it)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from ealdm_b200 import configs as CFG  # noqa: E402
from ealdm_b200.conditioner import UnetCond  # noqa: E402
from ealdm_b200.ops import Act  # noqa: E402
from oracle import autoencoder as OA  # noqa: E402  (checker only)
from oracle import conditioner as OC  # noqa: E402
from oracle import unet as OU  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def gold():
    return torch.load(os.path.join(GOLD, "conditioner.pt"), weights_only=False)


class _GivenEncoder(torch.nn.Module):
    """A first stage whose encoder output is the golden z (NHWC), so that the conditioner is tested alone."""

    def __init__(self, z_nchw):
        super().__init__()
        self.z = z_nchw

    def _eng(self, img):
        outer = self

        class E:
            def encoder_features(self, x):
                n, c, h, w = outer.z.shape
                buf = outer.z.permute(0, 2, 3, 1).reshape(n * h * w, c).contiguous().cuda()
                return Act(buf, n, h, w)
        return E()


def make_cond(convs):
    m = UnetCond(cond_args=dict(OC.COND_ARGS))
    m.load_state_dict(OC.synthetic_state_dict(), strict=True)
    m = m.cuda().eval()
    m.convs = convs
    return m


def images(G):
    return torch.rand(G["T"], 3, 256, 256, generator=torch.Generator().manual_seed(G["img_seed"])) * 2 - 1


def test_conditioner_state_dict_is_the_references():
    m = UnetCond(cond_args=dict(OC.COND_ARGS))
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == OC.param_shapes()


def test_conditioner_given_encoder_output_vs_reference_golden():
    G = gold()
    m = make_cond(_GivenEncoder(G["z"]))
    dummy = torch.zeros(G["T"], 3, 8, 8)
    ctx, inter = m((dummy, G["flow"], G["weather"], G["time"]), return_intermediates=True)
    assert ctx.shape == (G["T"], 4, 512) and ctx.dtype == torch.float32 and ctx.is_cuda
    for k in ("time_style", "flow_style", "weather_style"):
        err = rel_l2(inter[k], G[k])
        print(f"conditioner {k}: rel_l2 = {err:.3e}")
        assert err < 1e-5
    err = rel_l2(ctx, G["context_eval"])
    print(f"conditioner context (eval): rel_l2 = {err:.3e}")
    assert err < 1e-4
    # the 8-tuple the data loader yields (dataset_wlbl.py:567-570) goes the same way
    ctx8 = m((dummy, G["flow"], G["weather"], G["time"], 0, 0, 0, 0))
    assert torch.equal(ctx8, ctx)


def test_conditioner_batchnorm_training_mode_and_running_statistics():
    G = gold()
    m = make_cond(_GivenEncoder(G["z"]))
    m.conv_cat[1].train()
    bn_ref = torch.nn.BatchNorm2d(4)
    sd = OC.synthetic_state_dict()
    bn_ref.load_state_dict({k.split(".", 2)[2]: v for k, v in sd.items() if k.startswith("conv_cat.1.")})
    dummy = torch.zeros(G["T"], 3, 8, 8)
    ctx = m((dummy, G["flow"], G["weather"], G["time"]))
    err = rel_l2(ctx, G["context_bn_train"])
    print(f"conditioner context (BatchNorm on batch statistics): rel_l2 = {err:.3e}")
    assert err < 1e-4
    # running buffers: torch's update rule on the same conv output (from the oracle)
    import torch.nn.functional as F
    o = OC.unet_cond_forward(sd, G["z"], G["flow"], G["weather"], G["time"], bn_training=True)
    z = G["z"]
    cat = torch.cat((z, OC.adain(sd, "wadain", z, o["weather_style"]), OC.adain(sd, "fadain", z, o["flow_style"]),
                     OC.adain(sd, "tadain", z, o["time_style"])), dim=1)
    bn_ref.train()
    bn_ref(F.conv2d(cat, sd["conv_cat.0.weight"], sd["conv_cat.0.bias"], padding=1))
    assert rel_l2(m.conv_cat[1].running_mean, bn_ref.running_mean) < 1e-5
    assert rel_l2(m.conv_cat[1].running_var, bn_ref.running_var) < 1e-5
    assert int(m.conv_cat[1].num_batches_tracked) == int(bn_ref.num_batches_tracked)


def test_conditioner_lstm_recurrence_vs_reference_golden():
    G = gold()
    m = make_cond(None)
    out = m._lstm_mlp(m._pack()["w_mlp"], G["lstm_seq_in"].cuda())
    err = rel_l2(out, G["lstm_seq_out"])
    print(f"WeatherLSTM over 4 steps: rel_l2 = {err:.3e}")
    assert err < 1e-5


def test_conditioner_negative_branch_skips_the_styles():
    """mixed[-1] is None (the classifier-free 'negative' conditioning, ddpm.py:1322-1324): out_layer on the raw encoder map."""
    G = gold()
    m = make_cond(_GivenEncoder(G["z"]))
    dummy = torch.zeros(G["T"], 3, 8, 8)
    ctx = m((dummy, G["flow"], G["weather"], G["time"], 0, 0, 0, None))
    sd = OC.synthetic_state_dict()
    import torch.nn.functional as F
    h = F.relu(F.linear(G["z"].flatten(2), sd["out_layer.1.weight"], sd["out_layer.1.bias"]))
    ref = F.linear(h, sd["out_layer.4.weight"], sd["out_layer.4.bias"])
    assert rel_l2(ctx, ref) < 1e-4


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-4), ("bf16", 3e-2)])
def test_conditioner_end_to_end_with_vq_first_stage(mode, tol):
    """frames -> first-stage encoder (tcgen05 convs in bf16 mode) -> conditioner, wired as LatentDiffusion does."""
    from ealdm_b200.autoencoder import VQModelInterface
    G = gold()
    first = VQModelInterface(embed_dim=CFG.VQ_F8_EMBED_DIM, n_embed=CFG.VQ_F8_N_EMBED, ddconfig=dict(CFG.VQ_F8_DDCONFIG))
    first.load_state_dict(OU.synthetic_state_dict(
        OA.vq_param_shapes(CFG.VQ_F8_DDCONFIG, CFG.VQ_F8_EMBED_DIM, CFG.VQ_F8_N_EMBED), seed=4), strict=True)
    first = first.cuda().eval().set_compute_dtype(mode)
    m = make_cond(first)
    ctx, inter = m((images(G), G["flow"], G["weather"], G["time"]), return_intermediates=True)
    err = rel_l2(ctx, G["context_eval"])
    print(f"conditioner end to end, first stage {mode}: context rel_l2 = {err:.3e}")
    assert err < tol
    assert any(k.startswith("convs.encoder.") for k in m.state_dict())   # as in the reference's checkpoints


def test_conditioner_refuses_cpu():
    m = UnetCond(cond_args=dict(OC.COND_ARGS))
    G = gold()
    with pytest.raises(RuntimeError):
        m((torch.zeros(G["T"], 3, 8, 8), G["flow"], G["weather"], G["time"]))


def test_shipped_config_conditioner_to_sampler_pipeline():
    """The reference's shipped stdiff config with all four `target:` entries on the B200 modules, random-init:
    batch['mixed'] -> get_learned_conditioning (UnetCond on the first stage's encoder) -> the positive and the
    'negative' context (ddpm.py:1320-1326) -> 4 DDIM steps with classifier-free guidance -> decode_first_stage."""
    import yaml
    from ealdm_b200.ddim import DDIMSampler
    from ealdm_b200.util import instantiate_from_config
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = yaml.safe_load(open(os.path.join(root, "configs/latent-diffusion/stdiff_cin-ldm-vq-f8_b200.yaml")))["model"]
    torch.manual_seed(0)
    model = instantiate_from_config(cfg).cuda().eval()
    g = torch.Generator().manual_seed(5)
    B = 2
    mixed = [torch.rand(B, 3, 256, 256, generator=g) * 2 - 1, torch.rand(B, 1, 1, generator=g),
             torch.randn(B, 1, 16, generator=g), torch.rand(B, 1, generator=g), 0, 0, 0,
             torch.rand(B, 3, 256, 256, generator=g) * 2 - 1]
    c = model.get_learned_conditioning(mixed)
    neg = list(mixed)
    neg[0], neg[-1] = neg[-1], None
    uc = model.get_learned_conditioning(neg)
    assert c.shape == uc.shape == (B, 4, 512) and c.is_cuda and torch.isfinite(c).all() and torch.isfinite(uc).all()
    assert not torch.equal(c, uc)
    sampler = DDIMSampler(model)
    z, _ = sampler.sample(S=4, batch_size=B, shape=(4, 32, 32), conditioning=c, eta=0.0, verbose=False,
                          unconditional_guidance_scale=2.0, unconditional_conditioning=uc)
    x = model.decode_first_stage(z)
    assert x.shape == (B, 3, 256, 256) and torch.isfinite(x).all()


# ---- adjoint (the reference trains the conditioner with the UNet, ddpm.py:1409-1415) -----------------------------------
def _check_grad(name, got, want, tol=2e-4, floor=0.0):
    """`want` is a golden entry of oracle/gen_golden_cond_grads.py: the full tensor, or {norm, 4 projections, sample}.
    `floor`: gradients whose golden norm is below it are pure rounding noise (a bias in front of a batch-statistics
    BatchNorm has an exactly zero gradient) and are compared on that absolute scale."""
    from oracle.gen_golden_cond_grads import SAMPLE_STRIDE, direction
    got = got.detach().cpu()
    if "full" in want:
        ref = want["full"]
        if float(ref.abs().max()) == 0.0:
            assert float(got.abs().max()) == 0.0, name
            return 0.0
        err = float((got.double() - ref.double()).norm() / max(float(ref.double().norm()), floor))
    else:
        gd = got.double()
        scale = max(want["norm"], 1e-30)
        errs = [abs(float(gd.norm()) - want["norm"]) / scale]
        errs += [abs(float((gd * direction(name, k, got.shape).double()).sum()) - want["proj"][k]) / scale for k in range(4)]
        errs.append(float((got.reshape(-1)[::SAMPLE_STRIDE].double() - want["sample"].double()).norm()
                          / max(float(want["sample"].double().norm()), 1e-30)))
        err = max(errs)
    assert err < tol, (name, err)
    return err


@pytest.mark.parametrize("case", ["eval", "bn_train", "negative"])
def test_conditioner_gradients_vs_reference_autograd(case):
    """loss = sum(context * R): every own parameter's gradient from the hand-written adjoint (through torch.autograd:
    `_UnetCondFn`) against torch.autograd through the reference's own UnetCond (oracle/gen_golden_cond_grads.py)."""
    G = gold()
    GG = torch.load(os.path.join(GOLD, "conditioner_grads.pt"), weights_only=False)
    m = make_cond(_GivenEncoder(G["z"]))
    m.eval()                                  # Dropout off in every case (its mask is RNG plumbing)
    if case == "bn_train":
        m.conv_cat[1].train()
    T = G["T"]
    dummy = torch.zeros(T, 3, 8, 8)
    seed = GG["seeds"]["R_neg" if case == "negative" else "R"]
    R = torch.randn(T, 4, 512, generator=torch.Generator().manual_seed(seed)).cuda()
    mixed = ([dummy, G["flow"], G["weather"], G["time"], None, None, None, None] if case == "negative"
             else (dummy, G["flow"], G["weather"], G["time"]))
    for p in m.parameters():
        p.grad = None
    ctx = m(mixed)
    assert ctx.requires_grad
    want_ctx = GG["context_negative" if case == "negative" else f"context_{case}"]
    assert rel_l2(ctx, want_ctx) < 1e-4
    (ctx * R).sum().backward()
    worst = ("", 0.0)
    scale = max((float(v["full"].norm()) if "full" in v else v["norm"]) for v in GG[case].values())
    for name, p in m.named_parameters():
        if name.startswith("convs."):
            continue
        got = p.grad if p.grad is not None else torch.zeros_like(p)
        err = _check_grad(name, got, GG[case][name], floor=1e-4 * scale)
        if err > worst[1]:
            worst = (name, err)
    print(f"conditioner gradients [{case}]: worst relative error {worst[1]:.3e} ({worst[0]})")


def test_weather_lstm_backpropagation_through_time_vs_reference_autograd():
    """WeatherLSTM over a 4-step sequence (the shipped data has one step per frame): BPTT through ealdm_lstm_cell_bwd."""
    from ealdm_b200 import ops
    G = gold()
    GG = torch.load(os.path.join(GOLD, "conditioner_grads.pt"), weights_only=False)
    m = make_cond(_GivenEncoder(G["z"])).eval()
    seq = G["lstm_seq_in"].cuda()
    Rs = torch.randn(seq.shape[0] * seq.shape[1], 128, generator=torch.Generator().manual_seed(GG["seeds"]["R_seq"])).cuda()
    P = m._pack()
    sv = {}
    with torch.no_grad():
        out = m._lstm_mlp(P["w_mlp"], seq, sv)
        assert rel_l2(out, G["lstm_seq_out"]) < 1e-5
        grads = {n: torch.zeros_like(p, dtype=torch.float32) for n, p in m._own_named_parameters()}
        ws = ops.Workspace(seq.device)
        m._lstm_mlp_bwd("w_mlp", P["w_mlp"], sv, Rs, grads, ws, lambda *a, **k: m._linear_bwd(grads, ws, *a, **k))
    worst = max(_check_grad(n, grads[n], GG["lstm_seq"][n]) for n in GG["lstm_seq"])
    print(f"WeatherLSTM BPTT over 4 steps: worst relative error {worst:.3e}")
