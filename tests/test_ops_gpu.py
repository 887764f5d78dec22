"""Parity of every C-ABI operator against plain PyTorch fp32 on the same seeded inputs (GPU tests).

Tolerances: fp32 kernels 2e-5 relative L2 (summation order only); bf16 tensor-core kernels are
compared with the fp32 result computed from the SAME bf16-rounded inputs, 1e-2 relative L2 / output
rounding; the sampler kernels are bit-exact.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from ealdm_b200 import _lib as L  # noqa: E402
from ealdm_b200 import ops  # noqa: E402
from ealdm_b200.ops import Act, ConvIn  # noqa: E402

DEV = "cuda"


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def to_act(x_nchw, dtype, ld=None, c0=0):
    n, c, h, w = x_nchw.shape
    ld = ld or c
    buf = torch.zeros((n * h * w, ld), dtype=dtype, device=DEV)
    buf[:, c0:c0 + c] = x_nchw.permute(0, 2, 3, 1).reshape(-1, c).to(dtype)
    return Act(buf, n, h, w, c, c0)


def from_act(a: Act):
    return a.view2d().float().reshape(a.n, a.h, a.w, a.c).permute(0, 3, 1, 2).contiguous()


def pack_w(w, dtype):  # [O, I, kh, kw] -> [O, kh*kw*I]
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous().to(dtype)


def g(seed):
    return torch.Generator(device="cpu").manual_seed(seed)


def test_library_and_device():
    lib = L.load()
    assert lib.ealdm_abi_version() == 2
    assert lib.ealdm_device_check() == 0, lib.ealdm_last_error()


@pytest.mark.parametrize("dtype,impl", [(torch.float32, L.IMPL_SIMT), (torch.bfloat16, L.IMPL_SIMT),
                                        (torch.bfloat16, L.IMPL_TCGEN05)])
@pytest.mark.parametrize("M,K,N", [(256, 256, 256), (300, 512, 384), (128, 1024, 1024), (77, 64, 40),
                                   (1024, 256, 2048)])
def test_linear(dtype, impl, M, K, N):
    if impl == L.IMPL_TCGEN05 and N % 8:
        pytest.skip("ld_out alignment")
    x = torch.randn(M, K, generator=g(1)).to(DEV)
    w = (torch.randn(N, K, generator=g(2)) / math.sqrt(K)).to(DEV)
    b = torch.randn(N, generator=g(3)).to(DEV)
    r = torch.randn(M, N, generator=g(4)).to(DEV)
    xa = Act(x.to(dtype).contiguous(), 1, 1, M)
    ra = Act(r.to(dtype).contiguous(), 1, 1, M)
    out = Act.empty(1, 1, M, N, dtype, DEV)
    ops.linear(xa, w.to(dtype).contiguous(), out, bias=b, residual=ra, impl=impl)
    ref = F.linear(xa.buf.float(), w.to(dtype).float(), b) + ra.buf.float()
    tol = 2e-5 if dtype == torch.float32 else 6e-3
    assert rel_l2(out.buf.float(), ref) < tol


@pytest.mark.parametrize("dtype,impl", [(torch.float32, L.IMPL_SIMT), (torch.bfloat16, L.IMPL_TCGEN05)])
@pytest.mark.parametrize("n,c,h,w,co", [(2, 64, 32, 32, 128), (3, 128, 16, 16, 256), (5, 256, 8, 8, 64),
                                        (1, 64, 64, 64, 32), (2, 192, 8, 8, 4)])
def test_conv3x3_bias_rowvec(dtype, impl, n, c, h, w, co):
    if impl == L.IMPL_TCGEN05 and co % 8:
        pytest.skip("ld_out alignment handled by the host (padded out buffer)")
    x = torch.randn(n, c, h, w, generator=g(5)).to(DEV)
    wt = (torch.randn(co, c, 3, 3, generator=g(6)) / math.sqrt(9 * c)).to(DEV)
    b = torch.randn(co, generator=g(7)).to(DEV)
    emb = torch.randn(n, co + 16, generator=g(8)).to(DEV)  # rowvec with a column offset
    xa = to_act(x, dtype, ld=c + 8, c0=8)  # exercises pitch != channels
    out = Act.empty(n, h, w, co, dtype, DEV)
    ops.conv([ConvIn(xa, 3, 1, 1)], pack_w(wt, dtype), out, bias=b, rowvec=emb, rowvec_col0=16, impl=impl)
    xr = from_act(xa)
    ref = F.conv2d(xr, wt.to(dtype).float(), b, padding=1) + emb[:, 16:, None, None]
    tol = 2e-5 if dtype == torch.float32 else 6e-3
    assert rel_l2(from_act(out), ref) < tol


@pytest.mark.parametrize("dtype,impl", [(torch.float32, L.IMPL_SIMT), (torch.bfloat16, L.IMPL_SIMT),
                                        (torch.bfloat16, L.IMPL_TCGEN05)])
@pytest.mark.parametrize("pad", [1, 0])
def test_conv3x3_stride2(dtype, impl, pad):
    # pad=1: UNet Downsample (openaimodel.py:151); pad=0 with implicit bottom/right zero pad:
    # autoencoder Downsample, F.pad(x,(0,1,0,1)) + conv s2 p0 (model.py:74-76)
    n, c, h, w, co = 3, 128, 16, 16, 128
    x = torch.randn(n, c, h, w, generator=g(9)).to(DEV)
    wt = (torch.randn(co, c, 3, 3, generator=g(10)) / math.sqrt(9 * c)).to(DEV)
    b = torch.randn(co, generator=g(11)).to(DEV)
    xa = to_act(x, dtype)
    out = Act.empty(n, h // 2, w // 2, co, dtype, DEV)
    ops.conv([ConvIn(xa, 3, 2, pad)], pack_w(wt, dtype), out, bias=b, impl=impl)
    xr = from_act(xa)
    if pad == 1:
        ref = F.conv2d(xr, wt.to(dtype).float(), b, stride=2, padding=1)
    else:
        ref = F.conv2d(F.pad(xr, (0, 1, 0, 1)), wt.to(dtype).float(), b, stride=2, padding=0)
    tol = 2e-5 if dtype == torch.float32 else 6e-3
    assert rel_l2(from_act(out), ref) < tol


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_conv3x3_upsample_simt(dtype):
    n, c, h, w, co = 2, 64, 8, 8, 64
    x = torch.randn(n, c, h, w, generator=g(12)).to(DEV)
    wt = (torch.randn(co, c, 3, 3, generator=g(13)) / math.sqrt(9 * c)).to(DEV)
    xa = to_act(x, dtype)
    out = Act.empty(n, 2 * h, 2 * w, co, dtype, DEV)
    ops.conv([ConvIn(xa, 3, 1, 1, upsample=1)], pack_w(wt, dtype), out, impl=L.IMPL_SIMT)
    ref = F.conv2d(F.interpolate(from_act(xa), scale_factor=2, mode="nearest"), wt.to(dtype).float(), padding=1)
    tol = 2e-5 if dtype == torch.float32 else 6e-3
    assert rel_l2(from_act(out), ref) < tol
    up = Act.empty(n, 2 * h, 2 * w, c, dtype, DEV)
    ops.upsample_nearest2x(xa, up)
    assert torch.equal(from_act(up), F.interpolate(from_act(xa), scale_factor=2, mode="nearest"))


@pytest.mark.parametrize("n,c,h,w,co", [(2, 64, 8, 8, 64), (3, 128, 16, 16, 256), (4, 512, 8, 8, 1024),
                                        (5, 256, 16, 16, 320), (2, 64, 4, 8, 128)])
def test_conv3x3_upsample_as_four_phases_tcgen05(n, c, h, w, co):
    """Upsample (openaimodel.py:109-119) as four 2x2 output phases over the low-resolution input (`upsample_phases`):
    against F.interpolate(nearest) + conv2d on the same bf16-rounded input with the ORIGINAL 3x3 weights, including
    the fp32 output + bf16 shadow pair and the GroupNorm partial statistics the UNet asks for."""
    from ealdm_b200.packing import pack_upsample_phases
    x = torch.randn(n, c, h, w, generator=g(112)).to(DEV)
    wt = (torch.randn(co, c, 3, 3, generator=g(113)) / math.sqrt(9 * c)).to(DEV)
    b = torch.randn(co, generator=g(114)).to(DEV)
    xa = to_act(x, torch.bfloat16, ld=c + 64, c0=64)
    out = Act.empty(n, 2 * h, 2 * w, co, torch.float32, DEV).with_gn_partial()
    sh = Act.empty(n, 2 * h, 2 * w, co, torch.bfloat16, DEV)
    out.buf.fill_(float("nan"))
    wp = pack_upsample_phases(wt, torch.bfloat16)
    assert wp.shape == (co, 16 * c)
    ops.conv([ConvIn(xa, 3, 1, 1, upsample=1)], wp, out, bias=b, out2=sh, upsample_phases=True)
    up = F.interpolate(from_act(xa), scale_factor=2, mode="nearest")
    ref = F.conv2d(up, wt.to(torch.bfloat16).float(), b, padding=1)
    err = rel_l2(from_act(out), ref)
    print(f"upsample phases n={n} c={c} {h}x{w} -> co={co}: rel_l2 = {err:.3e}")
    assert err < 6e-3                      # pre-summed taps are rounded to bf16 once more than the 3x3 form
    assert torch.equal(sh.buf, out.buf.to(torch.bfloat16))
    # exactness of the phase algebra itself: the same kernel against conv2d with the phase weights un-summed again
    w4 = wp.float().reshape(co, 2, 2, 2, 2, c)      # [o, py, px, a, b, i]
    xr = F.pad(from_act(xa), (1, 1, 1, 1))
    ref2 = torch.empty_like(ref)
    for py in (0, 1):
        for px in (0, 1):
            k = w4[:, py, px].permute(0, 3, 1, 2).contiguous()            # [o, i, 2, 2]
            y = F.conv2d(xr[:, :, py:py + h + 1, px:px + w + 1], k, b)    # rows y + py - 1 + a (padded index + 1)
            ref2[:, :, py::2, px::2] = y
    assert rel_l2(from_act(out), ref2) < 2e-5
    if out.gp is not None and co % 128 == 0:   # GroupNorm (>= 4 channels per group) over the result equals torch's
        gam = torch.randn(co, generator=g(115)).to(DEV)
        bet = torch.randn(co, generator=g(116)).to(DEV)
        y = Act.empty(n, 2 * h, 2 * w, co, torch.bfloat16, DEV)
        ops.group_norm(out, gam, bet, 1e-5, y, silu=True)
        yr = F.silu(F.group_norm(from_act(out), 32, gam, bet, 1e-5))
        assert rel_l2(from_act(y), yr) < 6e-3


@pytest.mark.parametrize("dtype,impl", [(torch.float32, L.IMPL_SIMT), (torch.bfloat16, L.IMPL_TCGEN05)])
def test_conv_two_sources_skip_fused(dtype, impl):
    # ResBlock tail: conv3x3(h) + conv1x1(x) + biases in ONE accumulator (openaimodel.py:275)
    n, c1, c2, h, w, co = 2, 128, 192, 16, 16, 128
    hh = torch.randn(n, c1, h, w, generator=g(14)).to(DEV)
    x = torch.randn(n, c2, h, w, generator=g(15)).to(DEV)
    w3 = (torch.randn(co, c1, 3, 3, generator=g(16)) / math.sqrt(9 * c1)).to(DEV)
    w1 = (torch.randn(co, c2, 1, 1, generator=g(17)) / math.sqrt(c2)).to(DEV)
    b = torch.randn(co, generator=g(18)).to(DEV)
    ha, xa = to_act(hh, dtype), to_act(x, dtype)
    wcat = torch.cat([pack_w(w3, dtype), pack_w(w1, dtype)], dim=1).contiguous()
    out = Act.empty(n, h, w, co, dtype, DEV)
    ops.conv([ConvIn(ha, 3, 1, 1), ConvIn(xa, 1, 1, 0)], wcat, out, bias=b, impl=impl)
    ref = F.conv2d(from_act(ha), w3.to(dtype).float(), b, padding=1) + F.conv2d(from_act(xa), w1.to(dtype).float())
    tol = 2e-5 if dtype == torch.float32 else 6e-3
    assert rel_l2(from_act(out), ref) < tol


@pytest.mark.parametrize("dtype,impl", [(torch.float32, L.IMPL_SIMT), (torch.bfloat16, L.IMPL_SIMT),
                                        (torch.bfloat16, L.IMPL_TCGEN05)])
def test_geglu_linear(dtype, impl):
    from ealdm_b200.packing import geglu_interleave
    M, C = 384, 256
    x = torch.randn(M, C, generator=g(19)).to(DEV)
    w = (torch.randn(8 * C, C, generator=g(20)) / math.sqrt(C)).to(DEV)
    b = torch.randn(8 * C, generator=g(21)).to(DEV)
    xa = Act(x.to(dtype).contiguous(), 1, 1, M)
    wp, bp = geglu_interleave(w.to(dtype), b)
    out = Act.empty(1, 1, M, 4 * C, dtype, DEV)
    ops.linear(xa, wp, out, bias=bp, act=L.ACT_GEGLU, impl=impl)
    y = F.linear(xa.buf.float(), w.to(dtype).float(), b)
    val, gate = y.chunk(2, dim=-1)
    ref = val * F.gelu(gate)
    tol = 2e-5 if dtype == torch.float32 else 6e-3
    assert rel_l2(out.buf.float(), ref) < tol


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_silu_epilogue_and_f32_out(dtype):
    M, K, N = 128, 256, 1024
    x = torch.randn(M, K, generator=g(22)).to(DEV)
    w = (torch.randn(N, K, generator=g(23)) / math.sqrt(K)).to(DEV)
    b = torch.randn(N, generator=g(24)).to(DEV)
    xa = Act(x.to(dtype).contiguous(), 1, 1, M)
    out = Act.empty(1, 1, M, N, torch.float32, DEV)
    ops.linear(xa, w.to(dtype).contiguous(), out, bias=b, act=L.ACT_SILU)
    ref = F.silu(F.linear(xa.buf.float(), w.to(dtype).float(), b))
    assert rel_l2(out.buf, ref) < (2e-5 if dtype == torch.float32 else 2e-3)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("n,c,h,w", [(3, 256, 32, 32), (2, 768, 16, 16), (5, 1536, 8, 8), (2, 128, 64, 64),
                                     (130, 512, 8, 8)])
@pytest.mark.parametrize("silu", [True, False])
def test_group_norm(dtype, n, c, h, w, silu):
    x = (torch.randn(n, c, h, w, generator=g(25)) * 2 + 0.5).to(DEV)
    gamma = torch.randn(c, generator=g(26)).to(DEV)
    beta = torch.randn(c, generator=g(27)).to(DEV)
    xa = to_act(x, dtype, ld=c + 4, c0=4)
    out = Act.empty(n, h, w, c, dtype, DEV)
    ops.group_norm(xa, gamma, beta, 1e-5, out, silu=silu)
    out2 = Act.empty(n, h, w, c, dtype, DEV)
    ops.group_norm(xa, gamma, beta, 1e-5, out2, silu=silu)
    assert torch.equal(out.buf, out2.buf)  # no atomics: bit-reproducible
    ref = F.group_norm(from_act(xa), 32, gamma, beta, 1e-5)
    if silu:
        ref = F.silu(ref)
    assert rel_l2(from_act(out), ref) < (2e-5 if dtype == torch.float32 else 4e-3)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("rows,c", [(1000, 256), (513, 512), (64, 1024)])
def test_layer_norm(dtype, rows, c):
    x = (torch.randn(rows, c, generator=g(28)) * 3 - 1).to(DEV)
    gamma = torch.randn(c, generator=g(29)).to(DEV)
    beta = torch.randn(c, generator=g(30)).to(DEV)
    xa = Act(x.to(dtype).contiguous(), 1, 1, rows)
    out = Act.empty(1, 1, rows, c, dtype, DEV)
    ops.layer_norm(xa, gamma, beta, 1e-5, out)
    ref = F.layer_norm(xa.buf.float(), (c,), gamma, beta, 1e-5)
    assert rel_l2(out.buf.float(), ref) < (2e-5 if dtype == torch.float32 else 4e-3)


def _attn_ref(q, k, v, scale):  # [b, h, n, d]
    s = torch.einsum("bhid,bhjd->bhij", q, k) * scale
    return torch.einsum("bhij,bhjd->bhid", s.softmax(-1), v)


@pytest.mark.parametrize("dtype,impl", [(torch.float32, L.IMPL_SIMT), (torch.bfloat16, L.IMPL_SIMT),
                                        (torch.bfloat16, L.IMPL_AUTO)])
@pytest.mark.parametrize("b,h,n", [(2, 8, 1024), (3, 16, 256), (4, 32, 64), (1, 2, 200)])
def test_self_attention_packed_qkv(dtype, impl, b, h, n):
    d = 32
    C_ = h * d
    qkv = torch.randn(b * n, 3 * C_, generator=g(31)).to(DEV).to(dtype).contiguous()
    buf = Act(qkv, b, 1, n)
    out = Act.empty(b, 1, n, C_, dtype, DEV)
    ops.attention(buf.cols(0, C_), buf.cols(C_, C_), buf.cols(2 * C_, C_), out, batch=b, heads=h, head_dim=d,
                  n_q=n, n_kv=n, scale=d ** -0.5, impl=impl)
    f = qkv.float().reshape(b, n, 3, h, d)
    ref = _attn_ref(f[:, :, 0].transpose(1, 2), f[:, :, 1].transpose(1, 2), f[:, :, 2].transpose(1, 2), d ** -0.5)
    ref = ref.transpose(1, 2).reshape(b * n, C_)
    assert rel_l2(out.buf.float(), ref) < (2e-5 if dtype == torch.float32 else 1e-2)


@pytest.mark.parametrize("b,h,n", [(2, 8, 1024), (3, 16, 256), (1, 1, 128), (5, 4, 384)])
@pytest.mark.parametrize("ramp", [0.0, 6.0])
def test_self_attention_tcgen05_lazy_rescale_and_lse(b, h, n, ramp):
    """The tcgen05 / TMEM kernel, forced.  ramp > 0 makes the logits grow along the key axis (q.k up to ~ramp*30),
    so the row maximum keeps rising across key tiles and the lazy O rescaling in TMEM actually runs."""
    d = 32
    C_ = h * d
    qkv = torch.randn(b * n, 3 * C_, generator=g(35)).to(DEV)
    if ramp:
        x = qkv.reshape(b, n, 3, h, d)
        x[:, :, 0] = x[:, :, 0].abs() * 0.5 + 0.5                                   # q > 0
        x[:, :, 1] += (torch.arange(n, device=DEV).float() / n * ramp)[None, :, None, None]   # k drifts upwards
    qkv = qkv.to(torch.bfloat16).contiguous()
    buf = Act(qkv, b, 1, n)
    out = Act.empty(b, 1, n, C_, torch.bfloat16, DEV)
    lse = torch.empty(b, h, n, device=DEV)
    ops.attention(buf.cols(0, C_), buf.cols(C_, C_), buf.cols(2 * C_, C_), out, batch=b, heads=h, head_dim=d,
                  n_q=n, n_kv=n, scale=d ** -0.5, impl=L.IMPL_TCGEN05, lse=lse)
    f = qkv.float().reshape(b, n, 3, h, d)
    q, k, v = (f[:, :, i].transpose(1, 2) for i in range(3))
    ref = _attn_ref(q, k, v, d ** -0.5).transpose(1, 2).reshape(b * n, C_)
    assert rel_l2(out.buf.float(), ref) < 1e-2
    ref_lse = torch.logsumexp(torch.einsum("bhid,bhjd->bhij", q, k) * d ** -0.5, dim=-1) / math.log(2.0)
    assert float((lse - ref_lse).abs().max()) < 2e-3 * max(1.0, float(ref_lse.abs().max()))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_attention_legacy_interleaved_heads(dtype):
    # QKVAttentionLegacy (openaimodel.py:365): channels are [head][q|k|v][32]
    b, h, n, d = 2, 8, 256, 32
    qkv = torch.randn(b * n, 3 * h * d, generator=g(32)).to(DEV).to(dtype).contiguous()
    buf = Act(qkv, b, 1, n)
    out = Act.empty(b, 1, n, h * d, dtype, DEV)
    ops.attention(buf.cols(0, 3 * h * d - 2 * d), buf.cols(d, 3 * h * d - 2 * d), buf.cols(2 * d, 3 * h * d - 2 * d),
                  out, batch=b, heads=h, head_dim=d, n_q=n, n_kv=n, scale=d ** -0.5,
                  head_stride_q=3 * d, head_stride_kv=3 * d)
    f = qkv.float().reshape(b, n, h, 3, d)
    ref = _attn_ref(f[:, :, :, 0].transpose(1, 2), f[:, :, :, 1].transpose(1, 2), f[:, :, :, 2].transpose(1, 2),
                    d ** -0.5)
    ref = ref.transpose(1, 2).reshape(b * n, h * d)
    assert rel_l2(out.buf.float(), ref) < (2e-5 if dtype == torch.float32 else 1e-2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("nk", [4, 1, 7])
def test_cross_attention_small_kv(dtype, nk):
    b, h, n, d = 3, 16, 256, 32
    C_ = h * d
    q = torch.randn(b * n, C_, generator=g(33)).to(DEV).to(dtype).contiguous()
    kv = torch.randn(b * nk, 2 * C_, generator=g(34)).to(DEV).to(dtype).contiguous()
    qa, kva = Act(q, b, 1, n), Act(kv, b, 1, nk)
    out = Act.empty(b, 1, n, C_, dtype, DEV)
    ops.attention(qa, kva.cols(0, C_), kva.cols(C_, C_), out, batch=b, heads=h, head_dim=d, n_q=n, n_kv=nk,
                  scale=d ** -0.5)
    qf = q.float().reshape(b, n, h, d).transpose(1, 2)
    kf = kv.float()[:, :C_].reshape(b, nk, h, d).transpose(1, 2)
    vf = kv.float()[:, C_:].reshape(b, nk, h, d).transpose(1, 2)
    ref = _attn_ref(qf, kf, vf, d ** -0.5).transpose(1, 2).reshape(b * n, C_)
    assert rel_l2(out.buf.float(), ref) < (2e-5 if dtype == torch.float32 else 4e-3)


@pytest.mark.parametrize("n,hh,ww,C", [(3, 32, 32, 256), (2, 16, 16, 512), (2, 16, 8, 256), (1, 32, 16, 1024),
                                       (2, 16, 16, 768), (6, 8, 8, 1024), (5, 8, 8, 1024)])
def test_cross_attention_collapsed_onto_the_context(n, hh, ww, C):
    """CrossAttention (ldm/modules/attention.py:170-193) with a 4-token context, collapsed onto the context
    (packing.collapse_cross_attention + ealdm_conv wi_*): logits = x U_n^T with the softmax over the 4 keys in the
    epilogue, then out = P Zt_n + bias + residual, U_n / Zt_n gathered per image from ONE projection of the context.
    Checked against torch fp32 on the same bf16-rounded inputs and against the q / k / v kernels it replaces."""
    from ealdm_b200.packing import collapse_cross_attention
    heads, d, T, E = C // 32, 32, 4, 512
    M = n * hh * ww
    bf = torch.bfloat16
    x = torch.randn(M, C, generator=g(901)).to(DEV).to(bf)
    ctx = torch.randn(n * T, E, generator=g(902)).to(DEV).to(bf)
    wq = (torch.randn(C, C, generator=g(903)) / math.sqrt(C)).to(DEV)
    wk = (torch.randn(C, E, generator=g(904)) * (2.0 / math.sqrt(E))).to(DEV)   # logits with a spread of a few units
    wv = (torch.randn(C, E, generator=g(905)) / math.sqrt(E)).to(DEV)
    wo = (torch.randn(C, C, generator=g(906)) / math.sqrt(C)).to(DEV)
    bo = torch.randn(C, generator=g(907)).to(DEV)
    res = torch.randn(M, C, generator=g(908)).to(DEV)
    scale = d ** -0.5
    # torch fp32 reference
    q = (x.float() @ wq.t()).reshape(n, hh * ww, heads, d).transpose(1, 2)
    k = (ctx.float() @ wk.t()).reshape(n, T, heads, d).transpose(1, 2)
    v = (ctx.float() @ wv.t()).reshape(n, T, heads, d).transpose(1, 2)
    o = _attn_ref(q, k, v, scale).transpose(1, 2).reshape(M, C)
    want = o @ wo.t() + bo + res
    # collapsed path
    G, H = collapse_cross_attention(wq, wk, wv, wo, heads, scale, bf)
    gh = torch.cat([G, H], dim=0)
    xc = Act.empty(n, 1, T, gh.shape[0], bf, DEV)
    ops.linear(Act(ctx, n, 1, T), gh, xc)
    xa = Act(x, n, hh, ww)
    two = hh * ww == 64      # two 64-token images per tile: [probabilities | zeros] or [zeros | probabilities] per row
    pr = Act.empty(n, hh, ww, heads * T * (2 if two else 1), bf, DEV)
    ops.conv([ConvIn(xa)], xc.buf, pr, act=L.ACT_SOFTMAX4, wimg=(0, T, heads, C))
    pbuf = pr.buf.float()
    if two:
        odd = ((torch.arange(M, device=DEV) // 64) % 2 == 1)[:, None]
        halves = pbuf.reshape(M, 2, heads * T)
        assert (torch.where(odd, halves[:, 0], halves[:, 1]) == 0).all()     # the other image's half is exactly zero
        pbuf = torch.where(odd, halves[:, 1], halves[:, 0])
    psum = pbuf.reshape(M, heads, T).sum(-1)
    assert (psum - 1).abs().max() < 2e-2            # rows of probabilities (bf16-rounded)
    pref = torch.softmax(torch.einsum("bhqd,bhkd->bhqk", q, k) * scale, dim=-1).transpose(1, 2).reshape(M, heads * T)
    assert (pbuf - pref).abs().max() < 2.5e-2
    ra = Act(res, n, hh, ww)
    out = Act.empty(n, hh, ww, C, torch.float32, DEV)
    ops.conv([ConvIn(pr)], xc.buf, out, bias=bo, residual=ra, adjoint=True, wimg=(heads * C, T, heads, C))
    # the kernels it replaces
    qa = Act.empty(n, hh, ww, C, bf, DEV)
    ops.linear(xa, wq.to(bf).contiguous(), qa)
    kv = Act.empty(n, 1, T, 2 * C, bf, DEV)
    ops.linear(Act(ctx, n, 1, T), torch.cat([wk, wv]).to(bf).contiguous(), kv)
    oa = Act.empty(n, hh, ww, C, bf, DEV)
    ops.attention(qa, kv.cols(0, C), kv.cols(C, C), oa, batch=n, heads=heads, head_dim=d, n_q=hh * ww, n_kv=T,
                  scale=scale)
    out_k = Act.empty(n, hh, ww, C, torch.float32, DEV)
    ops.linear(oa, wo.to(bf).contiguous(), out_k, bias=bo, residual=ra)
    # errors of the attention term alone (the residual would mask them)
    e_new = rel_l2(out.buf - res - bo, want - res - bo)
    e_old = rel_l2(out_k.buf - res - bo, want - res - bo)
    print(f"collapsed cross-attention n={n} {hh}x{ww} C={C}: {e_new:.3e} (q/k/v kernels {e_old:.3e}) vs torch fp32")
    assert e_new < 1e-2 and e_new < 1.5 * e_old + 2e-3


def test_timestep_embedding_and_layout():
    t = torch.tensor([1, 501, 981, 0, 999], dtype=torch.int64, device=DEV)
    half = 128
    freqs = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32) / half)
    out = torch.empty(5, 256, dtype=torch.float32, device=DEV)
    ops.timestep_embedding(t, 256, freqs.to(DEV), out)
    args = t.cpu()[:, None].float() * freqs[None]
    ref = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    assert (out.cpu() - ref).abs().max() < 2e-6
    x = torch.randn(3, 4, 32, 32, generator=g(35)).to(DEV)
    a = Act.empty(3, 32, 32, 4, torch.float32, DEV)
    ops.nchw_to_nhwc(x, a)
    assert torch.equal(from_act(a), x)
    y = torch.empty_like(x)
    ops.nhwc_to_nchw(a, y)
    assert torch.equal(y, x)
    ab = Act.empty(3, 32, 32, 4, torch.bfloat16, DEV)
    ops.nchw_to_nhwc(x, ab)
    assert torch.equal(from_act(ab), x.bfloat16().float())
    c = Act.empty(3, 32, 32, 4, torch.float32, DEV)
    ops.copy2d(ab, c)
    assert torch.equal(c.buf, ab.buf.float())


def test_softmax_rows():
    x = torch.randn(300, 1024, generator=g(36)).to(DEV)
    a = Act(x.clone(), 1, 1, 300)
    ops.softmax_rows_(a, 0.125)
    assert rel_l2(a.buf, (x * 0.125).softmax(-1)) < 1e-5


@pytest.mark.parametrize("cfg", [True, False])
@pytest.mark.parametrize("eta", [0.0, 1.0])
def test_ddim_step_bit_exact(cfg, eta):
    # the reference's op sequence (ddim.py:173-203) evaluated with torch fp32 on the same device
    shape = (8, 4, 32, 32)
    x, eu, ec, nz = (torch.randn(shape, generator=g(40 + i)).to(DEV) for i in range(4))
    a_t, a_prev = torch.tensor(0.31234567, device=DEV), torch.tensor(0.4456789, device=DEV)
    sigma = eta * torch.sqrt((1 - a_prev) / (1 - a_t) * (1 - a_t / a_prev))
    s1 = torch.sqrt(1 - a_t)
    scale = 2.0
    e = eu + scale * (ec - eu) if cfg else ec
    pred_ref = (x - s1 * e) / a_t.sqrt()
    dir_ref = (1. - a_prev - sigma ** 2).sqrt() * e
    xprev_ref = a_prev.sqrt() * pred_ref + dir_ref + sigma * nz * 1.0
    xp, pr, eo = ops.ddim_step(x, ec, e_uncond=eu if cfg else None, noise=nz, cfg_scale=scale,
                               sqrt_one_minus_at=float(s1), sqrt_at=float(a_t.sqrt()),
                               sqrt_a_prev=float(a_prev.sqrt()),
                               dir_coef=float((1. - a_prev - sigma ** 2).sqrt()), sigma_t=float(sigma),
                               want_e=True)
    assert torch.equal(eo, e)
    assert torch.equal(pr, pred_ref)
    assert torch.equal(xp, xprev_ref)


def test_q_sample_bit_exact_and_cfg_mse():
    b = 16
    x0, nz, eu, ec = (torch.randn(b, 4, 32, 32, generator=g(50 + i)).to(DEV) for i in range(4))
    t = torch.randint(0, 1000, (b,), generator=g(54)).to(DEV)
    ac = torch.linspace(0.999, 0.001, 1000, device=DEV)
    sa, s1a = ac.sqrt(), (1 - ac).sqrt()
    out = ops.q_sample(x0, nz, t, sa, s1a)
    ref = sa[t].reshape(b, 1, 1, 1) * x0 + s1a[t].reshape(b, 1, 1, 1) * nz
    assert torch.equal(out, ref)
    ls = ops.cfg_mse(ec, nz, e_uncond=eu, cfg_scale=2.0)
    guided = eu + 2.0 * (ec - eu)
    ref = F.mse_loss(nz, guided, reduction="none").mean([1, 2, 3])
    assert rel_l2(ls, ref) < 1e-6
    ls2 = ops.cfg_mse(ec, nz)
    assert rel_l2(ls2, F.mse_loss(nz, ec, reduction="none").mean([1, 2, 3])) < 1e-6


def test_errors_are_reported():
    x = Act(torch.zeros(4, 6, device=DEV), 1, 1, 4)
    out = Act.empty(1, 1, 4, 6, torch.float32, DEV)
    with pytest.raises(L.EaldmError):
        ops.layer_norm(x, torch.ones(6, device=DEV), torch.zeros(6, device=DEV), 1e-5, out)  # c % 4 != 0


@pytest.mark.parametrize("n,c,h,w,co", [(2, 128, 32, 32, 256), (3, 256, 16, 16, 512), (5, 512, 8, 8, 1024),
                                        (2, 64, 64, 64, 128)])
@pytest.mark.parametrize("silu", [True, False])
def test_conv_epilogue_groupnorm_partials(n, c, h, w, co, silu):
    """The tcgen05 conv epilogue writes {sum, sum of squares} per (image, 32-pixel chunk, 8-channel octet); a
    GroupNorm given that buffer is one streaming pass and must equal GroupNorm of the stored fp32 output.  The
    output is a column window of a wider (concat) buffer whose other half is filled by the stand-alone kernel."""
    dtype = torch.bfloat16
    x = torch.randn(n, c, h, w, generator=g(61)).to(DEV)
    wt = (torch.randn(co, c, 3, 3, generator=g(62)) / math.sqrt(9 * c)).to(DEV)
    b = torch.randn(co, generator=g(63)).to(DEV)
    other = torch.randn(n, 64, h, w, generator=g(64)).to(DEV)
    cat = Act.empty(n, h, w, co + 64, torch.float32, DEV).with_gn_partial()
    assert cat.gp is not None
    left, right = cat.cols(0, co), cat.cols(co, 64)
    ops.conv([ConvIn(to_act(x, dtype), 3, 1, 1)], pack_w(wt, dtype), left, bias=b)
    right.view2d().copy_(other.permute(0, 2, 3, 1).reshape(-1, 64))
    ops.gn_partial(right)
    # the partials themselves
    full = from_act(cat)                                      # [n, co+64, h, w] fp32 as stored
    ref_p = full.reshape(n, (co + 64) // 8, 8, h * w // 32, 32)
    ref_sum = ref_p.sum(dim=(2, 4)).permute(0, 2, 1).reshape(-1, (co + 64) // 8)
    ref_sq = (ref_p ** 2).sum(dim=(2, 4)).permute(0, 2, 1).reshape(-1, (co + 64) // 8)
    assert rel_l2(cat.gp[..., 0], ref_sum) < 1e-5 and rel_l2(cat.gp[..., 1], ref_sq) < 1e-5
    # GroupNorm over the whole concat buffer (groups of (co+64)/32 channels) and over the left window alone
    for view in (cat, left):
        cc = view.c
        if (cc // 32) % 8:
            continue
        gamma = (1 + 0.2 * torch.randn(cc, generator=g(65))).to(DEV)
        beta = (0.2 * torch.randn(cc, generator=g(66))).to(DEV)
        y = Act.empty(n, h, w, cc, dtype, DEV)
        stats = torch.empty(n, 32, 2, device=DEV)
        ops.group_norm(view, gamma, beta, 1e-5, y, silu=silu, stats_out=stats)
        xr = from_act(view)
        ref = F.group_norm(xr, 32, gamma, beta, 1e-5)
        ref = F.silu(ref) if silu else ref
        assert rel_l2(from_act(y), ref) < 6e-3
        assert rel_l2(stats[..., 0], xr.reshape(n, 32, -1).mean(-1)) < 1e-4


@pytest.mark.parametrize("n,c,h,w,co,res", [(2, 128, 32, 32, 128, True), (3, 64, 16, 64, 128, False),
                                            (2, 128, 64, 64, 384, True)])
def test_conv_epilogue_groupnorm_partials_per_channel_quad(n, c, h, w, co, res):
    """gn_unit = 4: one {sum, sum of squares} entry per channel QUAD, for GroupNorm groups of 4 channels (the
    128-channel level of the autoencoder, model.py:116-141 Normalize = GroupNorm(32, c)): the GroupNorm fed by the
    quads equals GroupNorm of the stored fp32 output, with and without a residual in the producing epilogue."""
    dtype = torch.bfloat16
    x = torch.randn(n, c, h, w, generator=g(71)).to(DEV)
    wt = (torch.randn(co, c, 3, 3, generator=g(72)) / math.sqrt(9 * c)).to(DEV)
    b = torch.randn(co, generator=g(73)).to(DEV)
    out = Act.empty(n, h, w, co, torch.float32, DEV).with_gn_partial(unit=4)
    assert out.gp is not None and out.gp.shape[1] == co // 4
    r = Act(torch.randn(n * h * w, co, generator=g(74)).to(DEV), n, h, w) if res else None
    ops.conv([ConvIn(to_act(x, dtype), 3, 1, 1)], pack_w(wt, dtype), out, bias=b, residual=r)
    full = from_act(out)
    ref_p = full.reshape(n, co // 4, 4, h * w // 32, 32)
    ref_sum = ref_p.sum(dim=(2, 4)).permute(0, 2, 1).reshape(-1, co // 4)
    ref_sq = (ref_p ** 2).sum(dim=(2, 4)).permute(0, 2, 1).reshape(-1, co // 4)
    assert rel_l2(out.gp[..., 0], ref_sum) < 1e-5 and rel_l2(out.gp[..., 1], ref_sq) < 1e-5
    gamma = (1 + 0.2 * torch.randn(co, generator=g(75))).to(DEV)
    beta = (0.2 * torch.randn(co, generator=g(76))).to(DEV)
    y = Act.empty(n, h, w, co, dtype, DEV)
    ops.group_norm(out, gamma, beta, 1e-6, y, silu=True)
    ref = F.silu(F.group_norm(full, 32, gamma, beta, 1e-6))
    assert rel_l2(from_act(y), ref) < 6e-3


@pytest.mark.parametrize("M,K,res,geom", [(4096, 256, True, "linear"), (700, 256, False, "linear"),
                                          (2048, 1024, True, "linear"), (3 * 256, 64, True, "image")])
def test_layer_norm_applied_by_the_producing_epilogue(M, K, res, geom):
    """ealdm_conv ln_gamma / ln_beta: the GEMM that writes the fp32 stream x = a W^T + b (+ r) also writes
    LayerNorm(x) * gamma + beta as bf16 (two passes over the accumulator in TMEM, the two warps of a row exchanging
    their sums): equal to ealdm_layer_norm of the stored fp32 result up to bf16 rounding, ragged row counts, CTA pairs
    (K = 1024) and the (n, h, w) image geometry of the collapsed cross-attention included."""
    C = 256
    a = torch.randn(M, K, generator=g(801)).to(DEV).to(torch.bfloat16)
    W = (torch.randn(C, K, generator=g(802)) / math.sqrt(K)).to(DEV).to(torch.bfloat16)
    b = torch.randn(C, generator=g(803)).to(DEV)
    r = (torch.randn(M, C, generator=g(804)) * 2 + 0.7).to(DEV) if res else None
    gamma = (torch.rand(C, generator=g(805)) + 0.5).to(DEV)
    beta = (torch.randn(C, generator=g(806)) * 0.3).to(DEV)
    shape = (1, 1, M) if geom == "linear" else (3, 16, 16)
    aa = Act(a, *shape)
    x = Act.empty(*shape, C, torch.float32, DEV)
    y = Act.empty(*shape, C, torch.bfloat16, DEV)
    y.buf.fill_(float("nan"))
    ra = None if r is None else Act(r, *shape)
    ops.conv([ConvIn(aa)], W, x, bias=b, residual=ra, out2=y, ln_apply=(gamma, beta, 1e-5))
    want_x = a.float() @ W.float().t() + b + (r if res else 0)
    assert rel_l2(x.buf, want_x) < 1e-5 * 50
    ref = Act.empty(*shape, C, torch.bfloat16, DEV)
    ops.layer_norm(x, gamma, beta, 1e-5, ref)
    want = F.layer_norm(x.buf, (C,), gamma, beta, 1e-5)
    e_epi, e_ker = rel_l2(y.buf.float(), want), rel_l2(ref.buf.float(), want)
    print(f"LN in the epilogue M={M} K={K} res={res} {geom}: {e_epi:.3e} (LayerNorm kernel {e_ker:.3e}) vs torch fp32")
    assert e_epi < 4e-3 and e_epi < 1.2 * e_ker + 1e-4
    assert (y.buf.float() - ref.buf.float()).abs().max() < 0.07    # at most an ulp or two of bf16 apart


# ---- schedules of the tcgen05 conv (ealdm_tc_set_option): every setting must give the same numbers ----------------
class _tc_option:
    """Sets one schedule switch and forces the 256-column N tile (small problems would pick 128 by wave count and
    never reach the CTA-pair / wide-epilogue code)."""

    def __init__(self, opt, val, bn=256):
        self.opt, self.val, self.bn = opt, val, bn

    def __enter__(self):
        lib = L.load()
        self.prev = lib.ealdm_tc_set_option(self.opt, self.val)
        self.prev_bn = lib.ealdm_tc_set_option(L.TC_OPT_BN, self.bn)
        assert self.prev >= 0 and self.prev_bn >= 0

    def __exit__(self, *a):
        lib = L.load()
        lib.ealdm_tc_set_option(self.opt, self.prev)
        lib.ealdm_tc_set_option(L.TC_OPT_BN, self.prev_bn)


@pytest.mark.parametrize("n,c,h,w,co,res", [(4, 256, 16, 16, 256, False),    # 8 M tiles, K = 2304
                                            (6, 128, 16, 16, 512, True),     # 12 M tiles, 2 N tiles, fp32 residual
                                            (3, 128, 16, 16, 320, False),    # 6 M tiles, ragged N (320 = 256 + 64)
                                            (5, 64, 8, 8, 256, False)])      # 3 M tiles: odd -> pairs must decline
@pytest.mark.parametrize("bn", [256, 128])
def test_conv_cta_pairs_match_single_cta(n, c, h, w, co, res, bn):
    """cta_group::2 (256 x 256 tiles over a 2-CTA cluster, B split between the CTAs) against one CTA per tile:
    same k-block order, same fp32 accumulation -> bit-identical outputs; both against F.conv2d."""
    dtype = torch.bfloat16
    x = torch.randn(n, c, h, w, generator=g(70)).to(DEV)
    wt = (torch.randn(co, c, 3, 3, generator=g(71)) / math.sqrt(9 * c)).to(DEV)
    b = torch.randn(co, generator=g(72)).to(DEV)
    r = torch.randn(n, co, h, w, generator=g(73)).to(DEV)
    xa = to_act(x, dtype)
    ra = to_act(r, torch.float32) if res else None
    outs = []
    for mode in (0, 2):
        out = Act.empty(n, h, w, co, torch.float32, DEV)
        with _tc_option(L.TC_OPT_CTA2, mode, bn=bn):
            ops.conv([ConvIn(xa, 3, 1, 1)], pack_w(wt, dtype), out, bias=b, residual=ra, impl=L.IMPL_TCGEN05)
        torch.cuda.synchronize()
        outs.append(from_act(out))
    ref = F.conv2d(from_act(xa), wt.to(dtype).float(), b, padding=1) + (r if res else 0)
    assert rel_l2(outs[0], ref) < 2e-5 and rel_l2(outs[1], ref) < 2e-5   # fp32 out of bf16 operands: only sum order
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("n,c,h,w,co,feat", [
    (40, 64, 32, 32, 256, "only"),            # 320 tiles of one CTA on 148 SMs: images straddle the waves; groups of 8
    (40, 128, 32, 32, 256, "only+rowvec"),    # 160 super-tiles of CTA pairs (K = 1152)
    (24, 128, 16, 16, 512, "out2+res"),       # two N tiles, groups of 16, fp32 result + normalised copy
    (40, 128, 8, 8, 1024, "only+rowvec"),     # two images per tile, four N tiles, groups of 32
    (36, 128, 8, 8, 1024, "out2+skip1x1"),    # the ResBlock -> transformer case with the 1x1 skip accumulated
    (3, 64, 16, 16, 256, "only"),             # a few tiles only
    (5, 128, 8, 8, 1024, "out2+res"),         # odd image count: the last tile's second image does not exist
])
def test_group_norm_applied_by_the_producing_epilogue(n, c, h, w, co, feat):
    """ealdm_conv_args::gn_gamma: GroupNorm32 (+ SiLU) of the conv's result written by the conv's own epilogue -- the
    tiles of an image exchange their partial statistics through global memory and wait for each other.  Against
    F.group_norm of the un-fused conv's fp32 result (reference: GroupNorm32 + SiLU in front of a ResBlock's second conv,
    openaimodel.py:255-275) and against the stand-alone GroupNorm kernel; three runs (the counters re-arm themselves)."""
    dtype = torch.bfloat16
    x = torch.randn(n, c, h, w, generator=g(190)).to(DEV)
    wt = (torch.randn(co, c, 3, 3, generator=g(191)) / math.sqrt(9 * c)).to(DEV)
    b = torch.randn(co, generator=g(192)).to(DEV)
    gamma = (1.0 + 0.3 * torch.randn(co, generator=g(193))).to(DEV)
    beta = (0.3 * torch.randn(co, generator=g(194))).to(DEV)
    xa = to_act(x, dtype)
    kw, srcs, wp = {}, [ConvIn(xa, 3, 1, 1)], pack_w(wt, dtype)
    if "res" in feat:
        kw["residual"] = to_act(torch.randn(n, co, h, w, generator=g(195)).to(DEV), torch.float32)
    if "rowvec" in feat:
        kw["rowvec"] = 2.0 * torch.randn(n, co, generator=g(196)).to(DEV)     # per-image offsets: non-zero group means
    if "skip1x1" in feat:
        xs = to_act(torch.randn(n, 192, h, w, generator=g(197)).to(DEV), dtype)
        ws = (torch.randn(co, 192, 1, 1, generator=g(198)) / math.sqrt(192)).to(DEV)
        srcs.append(ConvIn(xs, 1, 1, 0))
        wp = torch.cat([wp, pack_w(ws, dtype)], dim=1).contiguous()
    only, silu = "only" in feat, "only" in feat
    plain = Act.empty(n, h, w, co, torch.float32, DEV).with_gn_partial()
    ops.conv(srcs, wp, plain, bias=b, impl=L.IMPL_TCGEN05, **kw)
    sep = Act.empty(n, h, w, co, dtype, DEV)
    ops.group_norm(plain, gamma, beta, 1e-5, sep, silu=silu)
    ref = F.group_norm(from_act(plain), 32, gamma, beta, 1e-5)
    ref = F.silu(ref) if silu else ref
    for _ in range(3):
        if only:
            y = Act.empty(n, h, w, co, dtype, DEV).with_gn_partial()
            ops.conv(srcs, wp, y, bias=b, gn_apply=(gamma, beta, 1e-5, 32, silu, True), **kw)
        else:
            f = Act.empty(n, h, w, co, torch.float32, DEV).with_gn_partial()
            y = Act.empty(n, h, w, co, dtype, DEV)
            ops.conv(srcs, wp, f, bias=b, out2=y, gn_apply=(gamma, beta, 1e-5, 32, silu, False), **kw)
            torch.cuda.synchronize()
            assert torch.equal(f.buf, plain.buf) and torch.equal(f.gp, plain.gp)
        torch.cuda.synchronize()
        got = from_act(y)
        assert rel_l2(got, ref) < 3e-3, rel_l2(got, ref)                       # bf16 rounding of the output
        assert rel_l2(got, from_act(sep)) < 2e-3, rel_l2(got, from_act(sep))   # the stand-alone kernel: an ulp here and there
        assert (got - from_act(sep)).abs().max() <= 0.04 * from_act(sep).abs().max()


@pytest.mark.parametrize("n,c,h,w,co,feat", [
    (128, 128, 8, 8, 1024, "rowvec"),        # 128 super-tiles on 74 clusters (the 8x8 level's shape): 1.73 tiles each
    (40, 128, 16, 16, 512, "res+out2+gn"),   # 80 super-tiles: remainder 6, every cluster parks and resumes
    (64, 128, 16, 16, 1024, "skip1x1"),      # 256 super-tiles: two whole waves + a dealt tail, two K segments
    (76, 128, 8, 8, 512, "res"),             # 19 x 2 = 38 < clusters: stream-K must decline
    (152, 128, 8, 8, 512, "plain"),          # 38 x 2 = 76 super-tiles: remainder 2
])
def test_conv_stream_k_is_bit_identical(n, c, h, w, co, feat):
    """Stream-K over the tail of the tile list (ealdm_tc_set_option STREAMK): the k-blocks of the last (clusters +
    remainder) super-tiles are dealt evenly; a split tile is begun by one cluster (accumulator parked in fp32) and
    resumed by the next from that accumulator, so the fp32 sums keep their k order.  Must equal the plain CTA-pair
    schedule bit for bit with every epilogue feature, run after run (the parking flags are re-armed by the reader)."""
    dtype = torch.bfloat16
    x = torch.randn(n, c, h, w, generator=g(170)).to(DEV)
    wt = (torch.randn(co, c, 3, 3, generator=g(171)) / math.sqrt(9 * c)).to(DEV)
    b = torch.randn(co, generator=g(172)).to(DEV)
    xa = to_act(x, dtype)
    kw, srcs, wp = {}, [ConvIn(xa, 3, 1, 1)], pack_w(wt, dtype)
    if "res" in feat:
        kw["residual"] = to_act(torch.randn(n, co, h, w, generator=g(173)).to(DEV), torch.float32)
    if "rowvec" in feat:
        kw["rowvec"] = torch.randn(n, co, generator=g(174)).to(DEV)
    if "skip1x1" in feat:
        xs = to_act(torch.randn(n, 192, h, w, generator=g(175)).to(DEV), dtype)
        ws = (torch.randn(co, 192, 1, 1, generator=g(176)) / math.sqrt(192)).to(DEV)
        srcs.append(ConvIn(xs, 1, 1, 0))
        wp = torch.cat([wp, pack_w(ws, dtype)], dim=1).contiguous()
    outs = []
    for mode in (0, 2, 2, 1):
        out = Act.empty(n, h, w, co, torch.float32, DEV)
        if "gn" in feat:
            out.with_gn_partial()
        o2 = Act.empty(n, h, w, co, dtype, DEV) if "out2" in feat else None
        with _tc_option(L.TC_OPT_STREAMK, mode, bn=256):
            ops.conv(srcs, wp, out, bias=b, out2=o2, impl=L.IMPL_TCGEN05, **kw)
        torch.cuda.synchronize()
        outs.append((from_act(out), None if o2 is None else from_act(o2), None if out.gp is None else out.gp.clone()))
    for o in outs[1:]:
        assert torch.equal(o[0], outs[0][0])
        assert (o[1] is None) or torch.equal(o[1], outs[0][1])
        assert (o[2] is None) or torch.equal(o[2], outs[0][2])
    if feat in ("plain", "rowvec"):
        ref = F.conv2d(from_act(xa), wt.to(dtype).float(), b, padding=1)
        if "rowvec" in feat:
            ref = ref + kw["rowvec"][:, :, None, None]
        assert rel_l2(outs[1][0], ref) < 2e-5


def test_linear_and_dgrad_stream_k_are_bit_identical():
    """The same for a K = 1024 linear layer with the fp32 residual stream (the 8x8 level's out-projection) and for the
    data-gradient mode (MN-major B operand, taps walked backwards)."""
    dtype = torch.bfloat16
    M, K, N = 8192 + 1024, 1024, 1024     # 36 x 4 = 144 super-tiles
    x = Act(torch.randn(M, K, generator=g(180)).to(DEV).to(dtype), 1, 1, M)
    w = (torch.randn(N, K, generator=g(181)) / math.sqrt(K)).to(DEV).to(dtype)
    b = torch.randn(N, generator=g(182)).to(DEV)
    res = Act(torch.randn(M, N, generator=g(183)).to(DEV), 1, 1, M)
    outs = []
    for mode in (0, 2, 2):
        out = Act.empty(1, 1, M, N, torch.float32, DEV)
        o2 = Act.empty(1, 1, M, N, dtype, DEV)
        with _tc_option(L.TC_OPT_STREAMK, mode, bn=256):
            ops.linear(x, w, out, bias=b, residual=res, out2=o2)
        torch.cuda.synchronize()
        outs.append((out.buf.clone(), o2.buf.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert torch.equal(outs[0][0], outs[2][0])
    n, c, h, w_, co = 160, 512, 8, 8, 128    # dX [160, 8, 8, 512]: 40 x 2 super-tiles of the adjoint GEMM, K = 1152
    wt = (torch.randn(co, c, 3, 3, generator=g(184)) / math.sqrt(9 * c)).to(DEV).to(dtype)
    dya = to_act(torch.randn(n, co, h, w_, generator=g(185)).to(DEV), dtype)
    packed = wt.permute(0, 2, 3, 1).reshape(co, -1).contiguous()
    douts = []
    for mode in (0, 2):
        dx = Act.empty(n, h, w_, c, torch.float32, DEV)
        with _tc_option(L.TC_OPT_STREAMK, mode, bn=256):
            ops.conv([ConvIn(dya, 3, 1, 1)], packed, dx, adjoint=True)
        torch.cuda.synchronize()
        douts.append(from_act(dx))
    assert torch.equal(douts[0], douts[1])


@pytest.mark.parametrize("M,K,N", [(4096, 1024, 512), (2048, 2048, 1024)])
def test_geglu_cta_pairs_match_single_cta(M, K, N):
    from ealdm_b200.packing import geglu_interleave
    dtype = torch.bfloat16
    x = torch.randn(M, K, generator=g(74)).to(DEV)
    w = (torch.randn(2 * N, K, generator=g(75)) / math.sqrt(K)).to(DEV)
    b = torch.randn(2 * N, generator=g(76)).to(DEV)
    xa = Act(x.to(dtype).contiguous(), 1, 1, M)
    wp, bp = geglu_interleave(w.to(dtype), b)
    outs = []
    for mode in (0, 2):
        out = Act.empty(1, 1, M, N, dtype, DEV)
        with _tc_option(L.TC_OPT_CTA2, mode):
            ops.linear(xa, wp, out, bias=bp, act=L.ACT_GEGLU, impl=L.IMPL_TCGEN05)
        torch.cuda.synchronize()
        outs.append(out.buf.float())
    y = F.linear(xa.buf.float(), w.to(dtype).float(), b)
    val, gate = y.chunk(2, dim=-1)
    assert rel_l2(outs[0], val * F.gelu(gate)) < 6e-3
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("M,K,N", [(2048, 256, 768), (1024, 256, 256), (640, 512, 320)])
def test_linear_wide_epilogue_matches_narrow(out_dtype, M, K, N):
    """128-byte-row epilogue passes (one fence / TMA store sequence per 64 or 128 columns) against the
    32-column units: identical arithmetic per element -> bit-identical outputs, ragged M and N included."""
    x = torch.randn(M, K, generator=g(77)).to(DEV).to(torch.bfloat16).contiguous()
    w = (torch.randn(N, K, generator=g(78)) / math.sqrt(K)).to(DEV).to(torch.bfloat16).contiguous()
    b = torch.randn(N, generator=g(79)).to(DEV)
    xa = Act(x, 1, 1, M)
    outs = []
    for mode in (0, 1):
        out = Act.empty(1, 1, M, N, out_dtype, DEV)
        with _tc_option(L.TC_OPT_WIDE, mode):
            ops.linear(xa, w, out, bias=b, impl=L.IMPL_TCGEN05)
        torch.cuda.synchronize()
        outs.append(out.buf.float())
    ref = F.linear(x.float(), w.float(), b)
    assert rel_l2(outs[0], ref) < (2e-5 if out_dtype == torch.float32 else 6e-3)
    assert torch.equal(outs[0], outs[1])


def test_geglu_wide_epilogue_matches_narrow():
    from ealdm_b200.packing import geglu_interleave
    M, C = 1024, 256
    x = torch.randn(M, C, generator=g(80)).to(DEV)
    w = (torch.randn(8 * C, C, generator=g(81)) / math.sqrt(C)).to(DEV)
    b = torch.randn(8 * C, generator=g(82)).to(DEV)
    xa = Act(x.to(torch.bfloat16).contiguous(), 1, 1, M)
    wp, bp = geglu_interleave(w.to(torch.bfloat16), b)
    outs = []
    for mode in (0, 1):
        out = Act.empty(1, 1, M, 4 * C, torch.bfloat16, DEV)
        with _tc_option(L.TC_OPT_WIDE, mode):
            ops.linear(xa, wp, out, bias=bp, act=L.ACT_GEGLU, impl=L.IMPL_TCGEN05)
        torch.cuda.synchronize()
        outs.append(out.buf.float())
    assert torch.equal(outs[0], outs[1])


def test_conv_cta_pairs_natural_tile_choice():
    """A problem large enough that the wave-count rule itself picks 256-column tiles and CTA pairs (no forcing):
    160 M tiles x 2 N tiles, K = 1152, checked against F.conv2d and against the single-CTA schedule."""
    dtype = torch.bfloat16
    n, c, h, w, co = 20, 128, 32, 32, 512
    x = torch.randn(n, c, h, w, generator=g(83)).to(DEV)
    wt = (torch.randn(co, c, 3, 3, generator=g(84)) / math.sqrt(9 * c)).to(DEV)
    b = torch.randn(co, generator=g(85)).to(DEV)
    xa = to_act(x, dtype)
    outs = []
    for mode in (0, 1):
        out = Act.empty(n, h, w, co, torch.float32, DEV)
        with _tc_option(L.TC_OPT_CTA2, mode, bn=0):
            ops.conv([ConvIn(xa, 3, 1, 1)], pack_w(wt, dtype), out, bias=b, impl=L.IMPL_TCGEN05)
        torch.cuda.synchronize()
        outs.append(from_act(out))
    ref = F.conv2d(from_act(xa), wt.to(dtype).float(), b, padding=1)
    assert rel_l2(outs[1], ref) < 2e-5
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("n,c,h,w,co", [(4, 256, 16, 16, 256), (2, 128, 16, 16, 192), (6, 512, 8, 8, 320)])
def test_dgrad_cta_pairs_match_single_cta(n, c, h, w, co):
    """weight_adjoint mode (data gradient from the forward-packed matrix, MN-major B operand) as CTA pairs: each CTA
    stages its half of the MN-major atoms; equals autograd's dX and the single-CTA schedule bit for bit."""
    dtype = torch.bfloat16
    x = torch.randn(n, c, h, w, generator=g(86)).to(DEV).to(dtype).float().requires_grad_(True)
    wt = (torch.randn(co, c, 3, 3, generator=g(87)) / math.sqrt(9 * c)).to(DEV).to(dtype).float()
    dy = torch.randn(n, co, h, w, generator=g(88)).to(DEV)
    dya = to_act(dy, dtype)
    F.conv2d(x, wt, padding=1).backward(from_act(dya))
    packed = wt.permute(0, 2, 3, 1).reshape(co, -1).to(dtype).contiguous()
    outs = []
    for mode in (0, 2):
        dx = Act.empty(n, h, w, c, torch.float32, DEV)
        with _tc_option(L.TC_OPT_CTA2, mode):
            ops.conv([ConvIn(dya, 3, 1, 1)], packed, dx, adjoint=True)
        torch.cuda.synchronize()
        outs.append(from_act(dx))
    assert rel_l2(outs[1], x.grad) < 8e-3
    assert torch.equal(outs[0], outs[1])


def _sweep_cases():
    import random
    rnd = random.Random(1234)
    cases = []
    for i in range(28):
        n = rnd.choice([1, 2, 3, 5, 6])
        hw = rnd.choice([(8, 8), (16, 16), (32, 32), (16, 8), (4, 32)])
        c = rnd.choice([64, 128, 192, 256])
        co = rnd.choice([64, 96, 160, 256, 320, 512])
        cases.append(dict(i=i, n=n, h=hw[0], w=hw[1], c=c, co=co, ksize=rnd.choice([1, 3, 3]),
                          stride=rnd.choice([1, 1, 1, 2]), skip_c=rnd.choice([0, 0, 64, 128]),
                          res=rnd.choice([None, "f32", "bf16"]), out_f32=rnd.choice([True, False]),
                          out2=rnd.choice([False, True]), rowvec=rnd.choice([False, True]),
                          gn=rnd.choice([False, True]), bias=rnd.choice([True, True, False])))
    return cases


@pytest.mark.parametrize("case", _sweep_cases(), ids=lambda c: f"case{c['i']}")
def test_conv_feature_sweep_all_schedules(case):
    """Seeded sweep over the conv epilogue's feature combinations (two sources, stride 2, fp32 / bf16 residual, fp32 /
    bf16 output, shadow output, per-image row vector, GroupNorm partial statistics, ragged N and M) run under every
    schedule -- default tile choice, forced 256-column tiles with and without CTA pairs, 128-column tiles as pairs --
    against F.conv2d; all schedules must agree bit for bit (same k-block order, same epilogue arithmetic)."""
    cs = dict(case)
    n, h, w, c, co, ks, stride = cs["n"], cs["h"], cs["w"], cs["c"], cs["co"], cs["ksize"], cs["stride"]
    if stride == 2 and (ks == 1 or h % 2 or w % 2):
        stride = 1
    ho, wo = h // stride, w // stride
    gn = cs["gn"] and cs["out_f32"] and co % 32 == 0 and (ho * wo) % 32 == 0
    out2 = cs["out2"] and cs["out_f32"]
    dt = torch.bfloat16
    x = torch.randn(n, c, h, w, generator=g(200 + cs["i"])).to(DEV)
    wt = (torch.randn(co, c, ks, ks, generator=g(300 + cs["i"])) / math.sqrt(ks * ks * c)).to(DEV)
    xa = to_act(x, dt)
    srcs = [ConvIn(xa, ks, stride, ks // 2)]
    wp = pack_w(wt, dt)
    ref = F.conv2d(from_act(xa), wt.to(dt).float(), stride=stride, padding=ks // 2)
    if cs["skip_c"] and stride == 1:
        xs = torch.randn(n, cs["skip_c"], h, w, generator=g(400 + cs["i"])).to(DEV)
        ws_ = (torch.randn(co, cs["skip_c"], 1, 1, generator=g(500 + cs["i"])) / math.sqrt(cs["skip_c"])).to(DEV)
        xsa = to_act(xs, dt)
        srcs.append(ConvIn(xsa, 1, 1, 0))
        wp = torch.cat([wp, pack_w(ws_, dt)], dim=1).contiguous()
        ref = ref + F.conv2d(from_act(xsa), ws_.to(dt).float())
    bias = torch.randn(co, generator=g(600 + cs["i"])).to(DEV) if cs["bias"] else None
    if bias is not None:
        ref = ref + bias[None, :, None, None]
    emb = None
    if cs["rowvec"] and co % 4 == 0:
        emb = torch.randn(n, co, generator=g(700 + cs["i"])).to(DEV)
        ref = ref + emb[:, :, None, None]
    ra = None
    if cs["res"] is not None:
        r = torch.randn(n, co, ho, wo, generator=g(800 + cs["i"])).to(DEV)
        ra = to_act(r, torch.float32 if cs["res"] == "f32" else dt)
        ref = ref + from_act(ra)
    odt = torch.float32 if cs["out_f32"] else dt
    results = []
    for bn, mode in ((0, 1), (256, 0), (256, 2), (128, 2)):
        out = Act.empty(n, ho, wo, co, odt, DEV)
        if gn:
            out.with_gn_partial()
        o2 = Act.empty(n, ho, wo, co, dt, DEV) if out2 else None
        with _tc_option(L.TC_OPT_CTA2, mode, bn=bn):
            ops.conv(srcs, wp, out, bias=bias, rowvec=emb, residual=ra, out2=o2, impl=L.IMPL_TCGEN05)
        torch.cuda.synchronize()
        results.append((from_act(out), None if o2 is None else from_act(o2),
                        out.gp.clone() if out.gp is not None else None))
    tol = 3e-5 if cs["out_f32"] else 6e-3
    assert rel_l2(results[0][0], ref) < tol
    for other in results[1:]:
        assert torch.equal(results[0][0], other[0])
        if out2:
            assert torch.equal(results[0][1], other[1]) and rel_l2(other[1], ref) < 6e-3
        if results[0][2] is not None:
            assert torch.equal(results[0][2], other[2])
    if results[0][2] is not None:   # the partial statistics really are the sums of the output
        o = results[0][0]
        sums = results[0][2].view(n, -1, co // 8, 2).sum(dim=1)           # [n, octet, {sum, sumsq}]
        ref_s = o.reshape(n, co // 8, 8, -1).sum(dim=(2, 3))
        ref_ss = (o * o).reshape(n, co // 8, 8, -1).sum(dim=(2, 3))
        assert rel_l2(sums[..., 0], ref_s) < 1e-4 and rel_l2(sums[..., 1], ref_ss) < 1e-4


@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("M", [128, 300, 1024, 40000])
def test_ff_geglu_fused_is_bit_identical_to_the_two_gemm_path(M, out_dtype):
    """ealdm_ff_geglu_fused (FF1 -> GEGLU -> FF2 -> + residual in one kernel, c = 256) against the unfused pair of
    ealdm_conv launches on the same interleaved weights (bit for bit: same k-block order, bias handling and GEGLU
    polynomial), and against torch fp32 on the bf16-rounded operands.  M = 40000 gives every persistent CTA several
    tiles (ring / TMEM buffer phases wrap); M = 300 ends in a partial tile."""
    from ealdm_b200.packing import geglu_interleave
    c, hid = 256, 1024
    x = torch.randn(M, c, generator=g(201)).to(DEV)
    w1 = (torch.randn(2 * hid, c, generator=g(202)) / math.sqrt(c)).to(DEV)
    b1 = (torch.randn(2 * hid, generator=g(203)) * 0.5).to(DEV)
    w2 = (torch.randn(c, hid, generator=g(204)) / math.sqrt(hid)).to(DEV)
    b2 = torch.randn(c, generator=g(205)).to(DEV)
    res = torch.randn(M, c, generator=g(206)).to(DEV)
    xa = Act(x.to(torch.bfloat16).contiguous(), 1, 1, M)
    ra = Act(res.contiguous(), 1, 1, M)
    w1i, b1i = geglu_interleave(w1.to(torch.bfloat16), b1)
    w2b = w2.to(torch.bfloat16).contiguous()
    out = Act.empty(1, 1, M, c, out_dtype, DEV)
    out.buf.fill_(float("nan"))
    ops.ff_geglu_fused(xa, w1i, b1i, w2b, b2, ra, out)
    gg = Act.empty(1, 1, M, hid, torch.bfloat16, DEV)
    ops.linear(xa, w1i, gg, bias=b1i, act=L.ACT_GEGLU)
    ref = Act.empty(1, 1, M, c, out_dtype, DEV)
    ops.linear(gg, w2b, ref, bias=b2, residual=ra)
    torch.cuda.synchronize()
    assert torch.equal(out.buf, ref.buf)
    xf = xa.buf.float()
    pre = F.linear(xf, w1.to(torch.bfloat16).float(), b1)
    h = (pre[:, :hid] * F.gelu(pre[:, hid:])).to(torch.bfloat16).float()
    want = F.linear(h, w2b.float(), b2) + res
    err = rel_l2(out.buf.float(), want)
    print(f"ff_geglu_fused M={M} out={out_dtype}: rel_l2 vs torch = {err:.3e}")
    assert err < 6e-3


def test_geglu_tanh_form_against_erf_form_and_torch():
    """The GEGLU epilogues evaluate GELU in its tanh form on MUFU.TANH by default (EALDM_TC_OPT_GELU_ERF = 0); the
    exact-erf form stays selectable.  Both against torch's exact GELU on the bf16-rounded operands, the fused
    FeedForward kernel in both forms against its unfused pair (bit for bit), and the two forms against each other."""
    from ealdm_b200.packing import geglu_interleave
    M, c, hid = 2048, 256, 1024
    x = torch.randn(M, c, generator=g(301)).to(DEV)
    w1 = (torch.randn(2 * hid, c, generator=g(302)) * (2.0 / math.sqrt(c))).to(DEV)     # gates spread over [-6, 6]
    b1 = torch.randn(2 * hid, generator=g(303)).to(DEV)
    w2 = (torch.randn(c, hid, generator=g(304)) / math.sqrt(hid)).to(DEV)
    b2 = torch.randn(c, generator=g(305)).to(DEV)
    res = torch.randn(M, c, generator=g(306)).to(DEV)
    xa = Act(x.to(torch.bfloat16).contiguous(), 1, 1, M)
    ra = Act(res.contiguous(), 1, 1, M)
    w1i, b1i = geglu_interleave(w1.to(torch.bfloat16), b1)
    w2b = w2.to(torch.bfloat16).contiguous()
    pre = F.linear(xa.buf.float(), w1.to(torch.bfloat16).float(), b1)
    want_h = pre[:, :hid] * F.gelu(pre[:, hid:])
    lib = L.load()
    got = {}
    prev = lib.ealdm_tc_set_option(L.TC_OPT_GELU_ERF, 0)
    try:
        for erf in (0, 1):
            lib.ealdm_tc_set_option(L.TC_OPT_GELU_ERF, erf)
            gg = Act.empty(1, 1, M, hid, torch.bfloat16, DEV)
            ops.linear(xa, w1i, gg, bias=b1i, act=L.ACT_GEGLU)
            two = Act.empty(1, 1, M, c, torch.float32, DEV)
            ops.linear(gg, w2b, two, bias=b2, residual=ra)
            one = Act.empty(1, 1, M, c, torch.float32, DEV)
            ops.ff_geglu_fused(xa, w1i, b1i, w2b, b2, ra, one)
            torch.cuda.synchronize()
            assert torch.equal(one.buf, two.buf)
            err = rel_l2(gg.buf.float(), want_h)
            dev_abs = float((gg.buf.float() - want_h).abs().max())
            print(f"GEGLU {'erf ' if erf else 'tanh'} form vs torch exact GELU: rel_l2 = {err:.3e}, max abs = {dev_abs:.3e}")
            assert err < 4e-3
            got[erf] = gg.buf.float()
    finally:
        lib.ealdm_tc_set_option(L.TC_OPT_GELU_ERF, prev)
    d = rel_l2(got[0], got[1])
    print(f"GEGLU tanh form vs erf form (both rounded to bf16): rel_l2 = {d:.3e}")
    assert d < 3e-3


@pytest.mark.parametrize("M,C,N,geglu", [(300, 256, 768, False), (4096, 256, 256, False), (2048, 512, 1536, False),
                                         (1024, 1024, 1024, False), (4096, 256, 2048, True), (1024, 512, 4096, True),
                                         (640, 1024, 8192, True)])
def test_layer_norm_folded_into_the_surrounding_gemms(M, C, N, geglu):
    """LayerNorm(x) W^T + b with the LayerNorm folded away (ealdm_conv ln_partial_out / ln_partial_in): a PRODUCER GEMM
    writes the fp32 stream x = y Wp^T + bp + r, its bf16 shadow and per-row {sum, sum of squares} partials; the CONSUMER
    GEMM multiplies the RAW shadow by W . gamma and corrects with the row statistics in its epilogue.  Checked against
    torch fp32 (LayerNorm of the SAME fp32 stream, operands rounded to bf16 like the unfused path) and against the
    unfused kernels (ealdm_layer_norm + ealdm_conv)."""
    from ealdm_b200.packing import geglu_interleave
    y = torch.randn(M, C, generator=g(401)).to(DEV)
    wp = (torch.randn(C, C, generator=g(402)) / math.sqrt(C)).to(DEV)
    bp = torch.randn(C, generator=g(403)).to(DEV)
    r = (torch.randn(M, C, generator=g(404)) * 2 + 0.5).to(DEV)          # a stream with a non-zero mean
    gamma = (torch.rand(C, generator=g(405)) + 0.5).to(DEV)
    beta = (torch.randn(C, generator=g(406)) * 0.3).to(DEV)
    W = (torch.randn(N, C, generator=g(407)) / math.sqrt(C)).to(DEV)
    b = torch.randn(N, generator=g(408)).to(DEV)
    ya = Act(y.to(torch.bfloat16).contiguous(), 1, 1, M)
    ra = Act(r.contiguous(), 1, 1, M)
    # producer: fp32 stream + bf16 shadow + row partials
    x = Act.empty(1, 1, M, C, torch.float32, DEV)
    xh = Act.empty(1, 1, M, C, torch.bfloat16, DEV)
    ops.linear(ya, wp.to(torch.bfloat16).contiguous(), x, bias=bp, residual=ra, out2=xh, ln_stats=True)
    assert x.ln is not None and x.ln.shape[0] == M and x.ln.shape[2] == 2
    xs = x.buf.float()
    s = x.ln.sum(1)
    assert rel_l2(s[:, 0], xs.sum(1)) < 1e-5 and rel_l2(s[:, 1], (xs * xs).sum(1)) < 1e-5
    assert torch.equal(xh.buf, x.buf.to(torch.bfloat16))
    # consumer
    wg = (W * gamma[None]).to(torch.bfloat16).contiguous()
    c1 = wg.float().sum(1).contiguous()
    c2 = (W @ beta + b).contiguous()
    act = L.ACT_GEGLU if geglu else L.ACT_NONE
    if geglu:
        wgi, c2i = geglu_interleave(wg, c2)
        _, c1i = geglu_interleave(wg, c1)
    else:
        wgi, c1i, c2i = wg, c1, c2
    out = Act.empty(1, 1, M, N // 2 if geglu else N, torch.bfloat16, DEV)
    ops.linear(xh, wgi, out, bias=c2i, act=act, ln=(x.ln, c1i, C, 1e-5))
    # unfused kernels on the same stream
    a = Act.empty(1, 1, M, C, torch.bfloat16, DEV)
    ops.layer_norm(x, gamma, beta, 1e-5, a)
    if geglu:
        wi, bi = geglu_interleave(W.to(torch.bfloat16), b)
    else:
        wi, bi = W.to(torch.bfloat16).contiguous(), b
    ref_k = Act.empty(1, 1, M, N // 2 if geglu else N, torch.bfloat16, DEV)
    ops.linear(a, wi, ref_k, bias=bi, act=act)
    # torch fp32
    pre = F.linear(F.layer_norm(xs, (C,), gamma, beta, 1e-5), W, b)
    want = pre[:, :N // 2] * F.gelu(pre[:, N // 2:]) if geglu else pre
    e_fold, e_unf = rel_l2(out.buf.float(), want), rel_l2(ref_k.buf.float(), want)
    print(f"LN fold M={M} C={C} N={N} geglu={geglu}: folded {e_fold:.3e}, unfused kernels {e_unf:.3e} (vs torch fp32)")
    assert e_fold < 6e-3 and e_fold < 1.5 * e_unf + 1e-3
