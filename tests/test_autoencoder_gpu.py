"""AutoencoderKL encode / decode on the CUDA path vs golden vectors made by the reference's
AutoencoderKL (oracle/gen_golden.py, 128x128 image / 16x16 latent, B=1).
fp32 mode: 1e-4 relative L2.  bf16 mode: encoder moments 1e-2 relative L2; for decoded images the
north star asks for a stated PSNR bound: PSNR >= 40 dB on the image mapped to [0, 1] (SURVEY.md
Appendix B).  The decoder's relative L2 is bounded by 1.5e-2: an IDEAL bf16-operand evaluation of this
49 M-parameter decoder (fp32 everywhere except GEMM inputs, emulated with the CPU oracle) already
differs from fp32 by 1.22e-2 with these random weights, so 1e-2 is not reachable by any bf16 path."""
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from ealdm_b200 import configs as CFG  # noqa: E402
from ealdm_b200.autoencoder import AutoencoderKL, DiagonalGaussianDistribution  # noqa: E402
from oracle import autoencoder as OA  # noqa: E402
from oracle import unet as OU  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = {"fp32": 1e-4, "bf16": 1e-2}


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


_ae = {}


def make_ae():
    if "ae" not in _ae:
        ae = AutoencoderKL(ddconfig=dict(CFG.AE_KL_F8_DDCONFIG), embed_dim=4)
        sd = OU.synthetic_state_dict(OA.autoencoder_kl_param_shapes(CFG.AE_KL_F8_DDCONFIG, 4), seed=3)
        ae.load_state_dict(sd, strict=True)
        _ae["ae"] = ae.cuda().eval()
    return _ae["ae"]


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_encode_vs_reference_golden(mode):
    G = torch.load(os.path.join(GOLD, "autoencoder_kl.pt"), weights_only=False)
    ae = make_ae().set_compute_dtype(mode)
    post = ae.encode(G["img"].cuda())
    assert isinstance(post, DiagonalGaussianDistribution)
    err = rel_l2(post.parameters, G["moments"])
    print(f"encode {mode}: moments rel_l2 = {err:.3e}")
    assert err < TOL[mode]
    assert rel_l2(post.mode(), G["mean"]) < TOL[mode]
    assert post.sample().shape == G["mean"].shape


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_decode_vs_reference_golden(mode):
    G = torch.load(os.path.join(GOLD, "autoencoder_kl.pt"), weights_only=False)
    ae = make_ae().set_compute_dtype(mode)
    dec = ae.decode(G["z"].cuda())
    err = rel_l2(dec, G["dec"])
    a = ((dec.cpu().clamp(-1, 1) + 1) / 2).double()
    b = ((G["dec"].clamp(-1, 1) + 1) / 2).double()
    mse = float(((a - b) ** 2).mean())
    psnr = 99.0 if mse == 0 else 10 * math.log10(1.0 / mse)
    print(f"decode {mode}: rel_l2 = {err:.3e}, PSNR = {psnr:.1f} dB")
    assert err < (TOL[mode] if mode == "fp32" else 1.5e-2)
    assert psnr >= 40.0
    assert torch.equal(dec, ae.decode(G["z"].cuda()))  # bit-reproducible


def test_decode_batch_of_three_matches_single():
    G = torch.load(os.path.join(GOLD, "autoencoder_kl.pt"), weights_only=False)
    ae = make_ae().set_compute_dtype("bf16")
    z = G["z"].cuda()
    z3 = torch.cat([z, z * 0.5, z])
    d3 = ae.decode(z3)
    d1 = ae.decode(z)
    assert rel_l2(d3[:1], d1) < 2e-3 and rel_l2(d3[2:], d1) < 2e-3


# ---- VQ first stage ---------------------------------------------------------------
def make_vq():
    if "vq" not in _ae:
        from ealdm_b200.autoencoder import VQModelInterface
        m = VQModelInterface(embed_dim=CFG.VQ_F8_EMBED_DIM, n_embed=CFG.VQ_F8_N_EMBED, ddconfig=dict(CFG.VQ_F8_DDCONFIG))
        sd = OU.synthetic_state_dict(OA.vq_param_shapes(CFG.VQ_F8_DDCONFIG, CFG.VQ_F8_EMBED_DIM, CFG.VQ_F8_N_EMBED), seed=4)
        sd["quantize.embedding.weight"] = torch.randn(CFG.VQ_F8_N_EMBED, CFG.VQ_F8_EMBED_DIM,
                                                      generator=torch.Generator().manual_seed(81)) * 1.2
        m.load_state_dict(sd, strict=True)
        _ae["vq"] = m.cuda().eval()
    return _ae["vq"]


def test_vq_nearest_indices_exact_vs_reference_golden():
    G = torch.load(os.path.join(GOLD, "vq_f8.pt"), weights_only=False)
    m = make_vq()
    zq, _, (_, _, idx) = m.quantize(G["h"].cuda())
    assert torch.equal(idx.cpu(), G["indices"])                                  # integer work: exact
    assert torch.equal(zq.cpu(), m.quantize.embedding.weight.detach().cpu()[G["indices"]].reshape(1, 32, 32, 4).permute(0, 3, 1, 2))
    # a batch: every pixel independent
    z = torch.randn(3, 4, 16, 16, generator=torch.Generator().manual_seed(5)).cuda()
    cb = m.quantize.embedding.weight.detach()
    d = ((z.permute(0, 2, 3, 1).reshape(-1, 1, 4).double() - cb.double()[None]) ** 2).sum(-1)
    _, _, (_, _, idx3) = m.quantize(z)
    assert float((d.argmin(1) == idx3).float().mean()) > 0.999


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_vq_decode_encode_vs_reference_golden(mode):
    G = torch.load(os.path.join(GOLD, "vq_f8.pt"), weights_only=False)
    m = make_vq().set_compute_dtype(mode)
    for key, kw in (("dec", {}), ("dec_noquant", {"force_not_quantize": True})):
        dec = m.decode(G["h"].cuda(), **kw)
        err = rel_l2(dec, G[key])
        a = ((dec.cpu().clamp(-1, 1) + 1) / 2).double()
        b = ((G[key].clamp(-1, 1) + 1) / 2).double()
        mse = float(((a - b) ** 2).mean())
        psnr = 99.0 if mse == 0 else 10 * math.log10(1.0 / mse)
        print(f"vq decode[{key}] {mode}: rel_l2 = {err:.3e}, PSNR = {psnr:.1f} dB")
        assert err < (TOL[mode] if mode == "fp32" else 2e-2) and psnr >= 40.0
    enc = m.encode(G["img"].cuda())
    assert rel_l2(enc, G["enc"]) < TOL[mode] * (1 if mode == "fp32" else 1.5)
