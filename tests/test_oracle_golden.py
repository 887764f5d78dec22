"""The CPU oracle against golden vectors produced by the reference's own modules
(oracle/gen_golden.py, run in the build container where /root/reference exists).  No GPU needed.

Tolerance: the oracle restates the same torch fp32 ops, so outputs agree to float rounding
(<= 1e-5 relative L2; schedules bit-exact)."""
import os

import pytest
import torch

from oracle import autoencoder as OA
from oracle import diffusion as OD
from oracle import unet as OU
from ealdm_b200 import configs as CFG

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def gold(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.fixture(scope="module")
def sd_uncond():
    return OU.synthetic_state_dict(OU.unet_param_shapes(CFG.UNET_UNCOND), seed=1)


@pytest.fixture(scope="module")
def sd_stdiff():
    return OU.synthetic_state_dict(OU.unet_param_shapes(CFG.UNET_STDIFF), seed=2)


def test_parameter_inventory_counts():
    # 394.98 M / 257.73 M parameters, 626 tensors for stdiff (SURVEY.md section 8b / BASELINE.md)
    s = OU.unet_param_shapes(CFG.UNET_STDIFF)
    u = OU.unet_param_shapes(CFG.UNET_UNCOND)
    n = lambda shapes: sum(int(torch.Size(sh).numel()) for _, sh in shapes)  # noqa: E731
    assert len(s) == 626
    assert abs(n(s) / 1e6 - 394.98) < 0.01 and abs(n(u) / 1e6 - 257.73) < 0.01
    a = OA.autoencoder_kl_param_shapes(CFG.AE_KL_F8_DDCONFIG, 4)
    enc = sum(int(torch.Size(sh).numel()) for k, sh in a if k.startswith("encoder."))
    dec = sum(int(torch.Size(sh).numel()) for k, sh in a if k.startswith("decoder."))
    assert abs(enc / 1e6 - 34.16) < 0.01 and abs(dec / 1e6 - 49.49) < 0.01
    assert gold("unet_stdiff_fwd.pt")["n_params"] == n(s)


def test_schedule_bit_exact():
    G = gold("schedule.pt")
    buf = OD.register_schedule(1000, CFG.DIFFUSION["linear_start"], CFG.DIFFUSION["linear_end"])
    for k, v in G["register"].items():
        assert torch.equal(buf[k], v), k
    for S in (10, 50):
        for eta in (0.0, 1.0):
            ref = G[f"S{S}_eta{eta}"]
            s = OD.make_ddim_schedule(buf["alphas_cumprod"], S, eta)
            assert torch.equal(torch.as_tensor(s["ddim_timesteps"]), ref["ddim_timesteps"])
            for k in ("ddim_alphas", "ddim_alphas_prev", "ddim_sigmas", "ddim_sqrt_one_minus_alphas"):
                mine = torch.as_tensor(__import__("numpy").asarray(s[k]), dtype=torch.float64)
                assert torch.equal(mine, ref[k]), (S, eta, k)
    # the scalars quoted in SURVEY.md a12
    s50 = OD.make_ddim_schedule(buf["alphas_cumprod"], 50, 1.0)
    assert list(s50["ddim_timesteps"][:3]) == [1, 21, 41] and s50["ddim_timesteps"][-1] == 981
    assert float(s50["ddim_alphas"][-1]) == 0.00020195601973682642
    assert float(s50["ddim_alphas_prev"][0]) == 0.9984999895095825
    assert float(s50["ddim_sigmas"][-1]) == 0.5611383211349679
    s10 = OD.make_ddim_schedule(buf["alphas_cumprod"], 10, 1.0)
    assert float(s10["ddim_alphas"][-1]) == 0.0008578449487686157
    assert float(s10["ddim_sigmas"][-1]) == 0.8883714867173162


def test_timestep_embedding():
    G = gold("schedule.pt")["timestep_embedding"]
    assert torch.equal(OU.timestep_embedding(G["t"], 256), G["emb"])


def test_unet_forward_uncond(sd_uncond):
    G = gold("unet_uncond_fwd.pt")
    with torch.no_grad():
        eps = OU.unet_forward(sd_uncond, CFG.UNET_UNCOND, G["x"], G["t"])
    assert float(G["eps"].std()) > 0.05          # not the degenerate all-zero reference init
    assert rel_l2(eps, G["eps"]) < 1e-5


def test_unet_forward_stdiff(sd_stdiff):
    G = gold("unet_stdiff_fwd.pt")
    with torch.no_grad():
        eps = OU.unet_forward(sd_stdiff, CFG.UNET_STDIFF, G["x"], G["t"], G["context"])
    assert float(G["eps"].std()) > 0.05
    assert rel_l2(eps, G["eps"]) < 1e-5


def test_unet_modules(sd_stdiff, sd_uncond):
    from oracle.gen_golden import golden_module_inputs
    G, I = gold("unet_modules.pt"), golden_module_inputs()
    import torch.nn.functional as F
    with torch.no_grad():
        y = OU.res_block(sd_stdiff, "input_blocks.4.0.", I["res_in4"]["x"], I["res_in4"]["emb"])
        assert rel_l2(y, G["res_in4"]) < 1e-5
        y = OU.spatial_transformer(sd_stdiff, "input_blocks.4.1.", I["st_in4"]["x"], I["st_in4"]["context"], 16)
        assert rel_l2(y, G["st_in4"]) < 1e-5
        y = F.conv2d(I["down_in3"]["x"], sd_stdiff["input_blocks.3.0.op.weight"],
                     sd_stdiff["input_blocks.3.0.op.bias"], stride=2, padding=1)
        assert rel_l2(y, G["down_in3"]) < 1e-5
        y = F.conv2d(F.interpolate(I["up_out2"]["x"], scale_factor=2, mode="nearest"),
                     sd_stdiff["output_blocks.2.2.conv.weight"], sd_stdiff["output_blocks.2.2.conv.bias"], padding=1)
        assert rel_l2(y, G["up_out2"]) < 1e-5
        y = OU.attention_block(sd_uncond, "input_blocks.4.1.", I["attnblock_in4"]["x"], 16)
        assert rel_l2(y, G["attnblock_in4"]) < 1e-5


def test_ddim_trajectory_config1(sd_uncond):
    """BASELINE.json configs[0]: uncond UNet, 10-step DDIM, batch 4, eta 0, fp32."""
    G = gold("ddim_traj.pt")["config1_uncond_B4_S10_eta0"]
    buf = OD.register_schedule(1000, CFG.DIFFUSION["linear_start"], CFG.DIFFUSION["linear_end"])
    apply_model = lambda x, t, c: OU.unet_forward(sd_uncond, CFG.UNET_UNCOND, x, t)  # noqa: E731
    with torch.no_grad():
        x0, trace = OD.ddim_sample(apply_model, buf["alphas_cumprod"], 10, G["x_T"], eta=0.0)
    for i, step in enumerate(trace):
        assert rel_l2(step["x_prev"], G["x_prev"][i]) < 2e-5, i
        assert rel_l2(step["pred_x0"], G["pred_x0"][i]) < 2e-5, i
    assert rel_l2(x0, G["samples"]) < 2e-5


def test_ddim_trajectory_stdiff_cfg_eta1(sd_stdiff):
    G = gold("ddim_traj.pt")["stdiff_B2_S10_eta1_cfg2"]
    buf = OD.register_schedule(1000, CFG.DIFFUSION["linear_start"], CFG.DIFFUSION["linear_end"])
    apply_model = lambda x, t, c: OU.unet_forward(sd_stdiff, CFG.UNET_STDIFF, x, t, c)  # noqa: E731
    with torch.no_grad():
        x0, trace = OD.ddim_sample(apply_model, buf["alphas_cumprod"], 10, G["x_T"], cond=G["cond"], eta=1.0,
                                   ugs=2.0, uc=G["uc"], noises=list(G["noise"]))
    for i, step in enumerate(trace):
        assert rel_l2(step["x_prev"], G["x_prev"][i]) < 5e-5, i
    assert rel_l2(x0, G["samples"]) < 5e-5


def test_plms_trajectory_uncond(sd_uncond):
    """PLMSSampler (plms.py): the oracle's restatement against the reference's own 10-step trajectory."""
    G = gold("plms_traj.pt")["uncond_B2_S10"]
    buf = OD.register_schedule(1000, CFG.DIFFUSION["linear_start"], CFG.DIFFUSION["linear_end"])
    apply_model = lambda x, t, c: OU.unet_forward(sd_uncond, CFG.UNET_UNCOND, x, t)  # noqa: E731
    with torch.no_grad():
        x0, trace = OD.plms_sample(apply_model, buf["alphas_cumprod"], 10, G["x_T"])
    for i, step in enumerate(trace):
        assert rel_l2(step["x_prev"], G["x_prev"][i]) < 2e-5, i
        assert rel_l2(step["pred_x0"], G["pred_x0"][i]) < 2e-5, i
    assert rel_l2(x0, G["samples"]) < 2e-5


def test_ddpm_ancestral_steps(sd_stdiff):
    """LatentDiffusion.p_sample_loop (timesteps 5..0): the oracle's p_sample restatement against the reference."""
    G = gold("ddpm_ancestral.pt")
    buf = OD.register_schedule(1000, CFG.DIFFUSION["linear_start"], CFG.DIFFUSION["linear_end"])
    apply_model = lambda x, t, c: OU.unet_forward(sd_stdiff, CFG.UNET_STDIFF, x, t, c)  # noqa: E731
    img = G["x_T"]
    with torch.no_grad():
        for k, i in enumerate(reversed(range(6))):
            t = torch.full((2,), i, dtype=torch.long)
            img, _ = OD.p_sample_ddpm(apply_model, buf, img, G["cond"], t, G["noise"][k], clip_denoised=G["clip_denoised"])
            assert rel_l2(img, G["imgs"][k]) < 2e-5, i
    assert rel_l2(img, G["out"]) < 2e-5


def test_p_losses_and_q_sample(sd_stdiff):
    G = gold("p_losses.pt")
    buf = OD.register_schedule(1000, CFG.DIFFUSION["linear_start"], CFG.DIFFUSION["linear_end"])
    assert torch.equal(OD.q_sample(buf, G["x0"], G["t"], G["noise"]), G["q_sample"])
    apply_model = lambda x, t, c: OU.unet_forward(sd_stdiff, CFG.UNET_STDIFF, x, t, c)  # noqa: E731
    with torch.no_grad():
        loss, d = OD.p_losses(apply_model, buf, G["x0"], G["cond2"], G["t"], G["noise"], ugs=2.0)
    assert abs(float(loss) - float(G["loss"])) <= 1e-5 * abs(float(G["loss"]))
    assert abs(float(d["loss_vlb"]) - float(G["loss_dict"]["val/loss_vlb"])) <= 1e-5 * abs(float(d["loss_vlb"]))


def test_autoencoder_kl():
    G = gold("autoencoder_kl.pt")
    dd = CFG.AE_KL_F8_DDCONFIG
    sd = OU.synthetic_state_dict(OA.autoencoder_kl_param_shapes(dd, 4), seed=3)
    with torch.no_grad():
        mom = OA.kl_encode_moments(sd, dd, G["img"])
        dec = OA.kl_decode(sd, dd, G["z"])
    assert rel_l2(mom, G["moments"]) < 1e-5
    mean, logvar, std = OA.gaussian_from_moments(mom)
    assert rel_l2(mean, G["mean"]) < 1e-5 and rel_l2(std, G["std"]) < 1e-5
    assert rel_l2(dec, G["dec"]) < 1e-5


def test_vq_first_stage_decode_and_quantizer():
    """VQ first stage (vq-f8: attention at 32x32, 16384 x 4 codebook): the oracle's Decoder / quantiser restatement
    against the reference VQModelInterface run with the restated taming quantiser (parity unpinned for the quantiser
    itself, see oracle/vq.py)."""
    from oracle import vq as OV
    G = gold("vq_f8.pt")
    shapes = OA.vq_param_shapes(CFG.VQ_F8_DDCONFIG, CFG.VQ_F8_EMBED_DIM, CFG.VQ_F8_N_EMBED)
    sd = OU.synthetic_state_dict(shapes, seed=4)
    sd["quantize.embedding.weight"] = torch.randn(CFG.VQ_F8_N_EMBED, CFG.VQ_F8_EMBED_DIM,
                                                  generator=torch.Generator().manual_seed(81)) * 1.2
    zq, idx = OV.vq_nearest(G["h"], sd["quantize.embedding.weight"])
    assert torch.equal(idx, G["indices"])
    # brute-force definition of the nearest code (float64)
    d = ((G["h"].permute(0, 2, 3, 1).reshape(-1, 1, 4).double() - sd["quantize.embedding.weight"].double()[None]) ** 2).sum(-1)
    assert float((d.argmin(1) == idx).float().mean()) > 0.999
    with torch.no_grad():
        dec = OA.vq_decode(sd, CFG.VQ_F8_DDCONFIG, G["h"])
        enc = OA.vq_encode(sd, CFG.VQ_F8_DDCONFIG, G["img"])
    assert rel_l2(dec, G["dec"]) < 2e-5
    assert rel_l2(enc, G["enc"]) < 2e-5


def test_conditioner_oracle_matches_reference_golden():
    """oracle/conditioner.py against tests/golden/conditioner.pt, produced by the reference's own UnetCond with the
    reference's VQModelInterface as `convs` (oracle/gen_golden_cond.py): context in eval mode and with BatchNorm on
    batch statistics, the three styles, the fourier features, and a 4-step LSTM recurrence."""
    from oracle import conditioner as OC
    G = gold("conditioner.pt")
    sd = OC.synthetic_state_dict()
    assert [k for k, _ in OC.param_shapes()] == list(sd.keys()) and len(sd) == 36
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())  # noqa: E731
    assert torch.equal(OC.fourier_features(G["time"]), G["fourier"])
    o = OC.unet_cond_forward(sd, G["z"], G["flow"], G["weather"], G["time"], bn_training=False)
    assert o["context"].shape == G["context_eval"].shape == (G["T"], 4, 512)
    assert rel(o["context"], G["context_eval"]) < 1e-5
    assert rel(o["time_style"], G["time_style"]) < 1e-6
    assert rel(o["flow_style"], G["flow_style"]) < 1e-6 and rel(o["weather_style"], G["weather_style"]) < 1e-6
    ob = OC.unet_cond_forward(sd, G["z"], G["flow"], G["weather"], G["time"], bn_training=True)
    assert rel(ob["context"], G["context_bn_train"]) < 1e-5
    assert rel(ob["context"], o["context"]) > 1e-3          # the two BatchNorm modes really differ
    assert rel(OC.lstm_mlp(sd, "w_mlp", G["lstm_seq_in"]), G["lstm_seq_out"]) < 1e-6
