"""Parity of every backward C-ABI operator against torch.autograd (fp32) on the same seeded inputs.

Tolerances: fp32 kernels 3e-5 relative L2 (summation order only); bf16 kernels are compared with the fp32
autograd result computed from the SAME bf16-rounded inputs: 8e-3 relative L2 (output rounding to bf16 where
the output is bf16, fp32 accumulation everywhere).
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
F32, BF16 = torch.float32, torch.bfloat16


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


def g(seed):
    return torch.Generator(device="cpu").manual_seed(seed)


def tol(dtype):
    return 3e-5 if dtype == F32 else 8e-3


# ---- weight gradients ------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype,impl", [(F32, L.IMPL_SIMT), (BF16, L.IMPL_SIMT), (BF16, L.IMPL_TCGEN05)])
@pytest.mark.parametrize("n,c,h,w,co,ks,stride,pad", [
    (2, 64, 32, 32, 128, 3, 1, 1),     # one channel block, two M halves
    (3, 256, 16, 16, 256, 3, 1, 1),    # BN = 256
    (5, 128, 8, 8, 320, 3, 1, 1),      # ragged M tile (320 = 2.5 x 128), batch not a power of two
    (2, 192, 8, 8, 64, 1, 1, 0),       # 1x1, c not a multiple of 128, half an M tile
    (3, 128, 16, 16, 128, 3, 2, 1),    # UNet Downsample
    (2, 64, 16, 16, 64, 3, 2, 0),      # autoencoder Downsample (implicit bottom/right zero pad)
    (2, 128, 12, 20, 72, 3, 1, 1),     # non power-of-two spatial extent, ragged co
    (4, 256, 32, 32, 8, 3, 1, 1),      # UNet out conv (4 channels padded to 8)
])
def test_conv_wgrad(dtype, impl, n, c, h, w, co, ks, stride, pad):
    x = torch.randn(n, c, h, w, generator=g(1)).to(DEV)
    if stride == 2:
        ho, wo = h // 2, w // 2
    else:
        ho, wo = h, w
    dy = torch.randn(n, co, ho, wo, generator=g(2)).to(DEV)
    xa = to_act(x, dtype, ld=c + 8, c0=8)
    dya = to_act(dy, dtype)
    ws = ops.Workspace(DEV)
    # reference: autograd of conv2d on the bf16-rounded operands
    xr = from_act(xa)
    wt = torch.zeros(co, c, ks, ks, device=DEV, requires_grad=True)
    if stride == 2 and pad == 0:
        y = F.conv2d(F.pad(xr, (0, 1, 0, 1)), wt, stride=2)
    else:
        y = F.conv2d(xr, wt, stride=stride, padding=pad)
    y.backward(from_act(dya))
    ref = wt.grad
    # OIHW, accumulate on top of a non-zero gradient
    base = torch.randn(co, c, ks, ks, generator=g(3)).to(DEV)
    dw = base.clone()
    ops.conv_wgrad(xa, dya, dw.view(co, -1), ws, ksize=ks, stride=stride, pad=pad, layout=L.WGRAD_OIHW,
                   accumulate=True, impl=impl)
    assert rel_l2(dw - base, ref) < tol(dtype)
    # packed K-major into a column window, overwrite
    kk = ks * ks * c
    dwp = torch.full((co, kk + 24), 7.0, device=DEV)
    ops.conv_wgrad(xa, dya, dwp, ws, ksize=ks, stride=stride, pad=pad, col0=16, layout=L.WGRAD_PACKED,
                   accumulate=False, impl=impl)
    refp = ref.permute(0, 2, 3, 1).reshape(co, kk)
    assert rel_l2(dwp[:, 16:16 + kk], refp) < tol(dtype)
    assert bool((dwp[:, :16] == 7.0).all()) and bool((dwp[:, 16 + kk:] == 7.0).all())


@pytest.mark.parametrize("dtype,impl", [(F32, L.IMPL_SIMT), (BF16, L.IMPL_TCGEN05)])
@pytest.mark.parametrize("M,K,N", [(2048, 256, 768), (300, 512, 384), (64, 1024, 1024), (8192, 256, 2048),
                                   (8, 512, 2048)])
def test_linear_wgrad(dtype, impl, M, K, N):
    x = torch.randn(M, K, generator=g(4)).to(DEV).to(dtype)
    dy = torch.randn(M, N, generator=g(5)).to(DEV).to(dtype)
    dw = torch.zeros(N, K, device=DEV)
    ops.linear_wgrad(Act(x, 1, 1, M), Act(dy, 1, 1, M), dw, ops.Workspace(DEV), accumulate=False, impl=impl)
    ref = dy.float().t() @ x.float()
    assert rel_l2(dw, ref) < tol(dtype)


def test_dgrad_is_conv_with_flipped_weights():
    """The data gradient of conv3x3 (stride 1, pad 1) is ealdm_conv with W[ci, (2-kh, 2-kw), co]; of the
    stride-2 conv, the same after zero insertion; of nearest-2x + conv, followed by 2x2 sum pooling."""
    n, c, h, w, co = 2, 128, 16, 16, 192
    dtype = BF16
    x = torch.randn(n, c, h, w, generator=g(6)).to(DEV).to(dtype).float().requires_grad_(True)
    wt = (torch.randn(co, c, 3, 3, generator=g(7)) / math.sqrt(9 * c)).to(DEV).to(dtype).float()
    wd = wt.flip(2, 3).permute(1, 2, 3, 0).reshape(c, 9 * co).contiguous().to(dtype)
    # stride 1
    dy = torch.randn(n, co, h, w, generator=g(8)).to(DEV)
    dya = to_act(dy, dtype)
    F.conv2d(x, wt, padding=1).backward(from_act(dya))
    dx = Act.empty(n, h, w, c, F32, DEV)
    ops.conv([ConvIn(dya, 3, 1, 1)], wd, dx)
    assert rel_l2(from_act(dx), x.grad) < 8e-3
    # stride 2
    x.grad = None
    dy2 = torch.randn(n, co, h // 2, w // 2, generator=g(9)).to(DEV)
    dy2a = to_act(dy2, dtype)
    F.conv2d(x, wt, stride=2, padding=1).backward(from_act(dy2a))
    z = ops.zero_insert2x(dy2a, Act.empty(n, h, w, co, dtype, DEV))
    ops.conv([ConvIn(z, 3, 1, 1)], wd, dx)
    assert rel_l2(from_act(dx), x.grad) < 8e-3
    # nearest-2x upsampling then conv
    x.grad = None
    dy3 = torch.randn(n, co, 2 * h, 2 * w, generator=g(10)).to(DEV)
    dy3a = to_act(dy3, dtype)
    F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), wt, padding=1).backward(from_act(dy3a))
    dup = Act.empty(n, 2 * h, 2 * w, c, dtype, DEV)
    ops.conv([ConvIn(dy3a, 3, 1, 1)], wd, dup)
    add = torch.randn(n * h * w, c, generator=g(11)).to(DEV)
    dx2 = Act.empty(n, h, w, c, dtype, DEV)
    ops.sumpool2x2(dup, dx, add=Act(add, n, h, w), dx2=dx2)
    ref = x.grad + add.reshape(n, h, w, c).permute(0, 3, 1, 2)
    assert rel_l2(from_act(dx), ref) < 8e-3
    assert rel_l2(from_act(dx2), ref) < 8e-3


@pytest.mark.parametrize("n,c,h,w,co", [(2, 128, 16, 16, 192), (3, 256, 8, 8, 512), (1, 64, 32, 32, 64)])
def test_dgrad_adjoint_mode_reads_forward_weights(n, c, h, w, co):
    """weight_adjoint: the data gradient straight from the FORWARD-packed matrix (MN-major B operand), including a
    column window of a fused 3x3 + 1x1-skip matrix; equals autograd's dX of conv2d."""
    dtype = BF16
    skip_c = 128
    x = torch.randn(n, c, h, w, generator=g(40)).to(DEV).to(dtype).float().requires_grad_(True)
    xs = torch.randn(n, skip_c, h, w, generator=g(41)).to(DEV).to(dtype).float().requires_grad_(True)
    wt = (torch.randn(co, c, 3, 3, generator=g(42)) / math.sqrt(9 * c)).to(DEV).to(dtype).float()
    ws = (torch.randn(co, skip_c, 1, 1, generator=g(43)) / math.sqrt(skip_c)).to(DEV).to(dtype).float()
    dy = torch.randn(n, co, h, w, generator=g(44)).to(DEV)
    dya = to_act(dy, dtype)
    (F.conv2d(x, wt, padding=1) + F.conv2d(xs, ws)).backward(from_act(dya))
    # the forward engine's fused matrix: [co, 9*c | skip_c]
    fused = torch.cat([wt.permute(0, 2, 3, 1).reshape(co, -1), ws.reshape(co, skip_c)], dim=1).to(dtype).contiguous()
    dx = Act.empty(n, h, w, c, F32, DEV)
    ops.conv([ConvIn(dya, 3, 1, 1)], fused[:, :9 * c], dx, adjoint=True)
    assert rel_l2(from_act(dx), x.grad) < 8e-3
    res = torch.randn(n * h * w, skip_c, generator=g(45)).to(DEV)
    dxs = Act.empty(n, h, w, skip_c, F32, DEV)
    dxs2 = Act.empty(n, h, w, skip_c, dtype, DEV)
    ops.conv([ConvIn(dya, 1, 1, 0)], fused[:, 9 * c:], dxs, adjoint=True, residual=Act(res, n, h, w), out2=dxs2)
    ref = xs.grad + res.reshape(n, h, w, skip_c).permute(0, 3, 1, 2)
    assert rel_l2(from_act(dxs), ref) < 8e-3 and rel_l2(from_act(dxs2), ref) < 8e-3
    # a Linear layer: dX = dY W
    M, K, N = 640, 256, 512
    wl = (torch.randn(N, K, generator=g(46)) / math.sqrt(K)).to(DEV).to(dtype)
    dyl = torch.randn(M, N, generator=g(47)).to(DEV).to(dtype)
    dxl = Act.empty(1, 1, M, K, F32, DEV)
    ops.linear(Act(dyl, 1, 1, M), wl, dxl, adjoint=True)
    assert rel_l2(dxl.buf, dyl.float() @ wl.float()) < 8e-3


# ---- normalisation -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype,x_f32", [(F32, False), (BF16, True), (BF16, False)])
@pytest.mark.parametrize("silu", [True, False])
@pytest.mark.parametrize("n,c,h,w", [(2, 256, 32, 32), (3, 512, 16, 16), (5, 1024, 8, 8), (2, 2048, 8, 8),
                                     (1, 128, 20, 12)])
def test_group_norm_bwd(dtype, x_f32, silu, n, c, h, w):
    eps = 1e-5
    x = (torch.randn(n, c, h, w, generator=g(12)) * 1.5 + 0.3).to(DEV)
    gamma = (1 + 0.2 * torch.randn(c, generator=g(13))).to(DEV)
    beta = (0.2 * torch.randn(c, generator=g(14))).to(DEV)
    dy = torch.randn(n, c, h, w, generator=g(15)).to(DEV)
    add = torch.randn(n, c, h, w, generator=g(16)).to(DEV)
    add2 = torch.randn(n, c + 32, h, w, generator=g(17)).to(DEV)
    xdt = F32 if (dtype == F32 or x_f32) else BF16
    xa = to_act(x, xdt)
    dya = to_act(dy, dtype)
    # forward through the library to get the saved statistics
    y = Act.empty(n, h, w, c, dtype, DEV)
    stats = torch.empty(n, 32, 2, device=DEV)
    ops.group_norm(xa, gamma, beta, eps, y, silu=silu, stats_out=stats)
    xr = from_act(xa).requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = F.group_norm(xr, 32, gr, br, eps)
    if silu:
        yr = F.silu(yr)
    yr.backward(from_act(dya))
    mean_ref = xr.detach().reshape(n, 32, -1).mean(-1)
    assert rel_l2(stats[..., 0], mean_ref) < 1e-4
    ws = ops.Workspace(DEV)
    dx = Act.empty(n, h, w, c, F32, DEV)
    dx2 = Act.empty(n, h, w, c, dtype, DEV) if dtype != F32 else None
    dg0, db0 = torch.randn(c, generator=g(18)).to(DEV), torch.randn(c, generator=g(19)).to(DEV)
    dg, db = dg0.clone(), db0.clone()
    adda = to_act(add, F32)
    add2a = to_act(add2, F32).cols(16, c)
    ops.group_norm_bwd(xa, dya, stats, gamma, beta, dx, ws, silu=silu, add=adda, add2=add2a, dx2=dx2,
                       dgamma=dg, dbeta=db)
    ref = xr.grad + add + add2[:, 16:16 + c]
    assert rel_l2(from_act(dx), ref) < 3e-5 * (1 if dtype == F32 else 10)
    if dx2 is not None:
        assert rel_l2(from_act(dx2), ref) < 8e-3
    assert rel_l2(dg - dg0, gr.grad) < 1e-4
    assert rel_l2(db - db0, br.grad) < 1e-4


@pytest.mark.parametrize("dtype", [F32, BF16])
@pytest.mark.parametrize("rows,c", [(2048, 256), (515, 512), (64, 1024), (40000, 256), (3, 128)])
def test_layer_norm_bwd(dtype, rows, c):
    eps = 1e-5
    x = (torch.randn(rows, c, generator=g(20)) * 2 + 0.5).to(DEV)
    gamma = (1 + 0.2 * torch.randn(c, generator=g(21))).to(DEV)
    beta = (0.2 * torch.randn(c, generator=g(22))).to(DEV)
    dy = torch.randn(rows, c, generator=g(23)).to(DEV).to(dtype)
    add = torch.randn(rows, c, generator=g(24)).to(DEV)
    xr = x.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    F.layer_norm(xr, (c,), gr, br, eps).backward(dy.float())
    dx = Act.empty(1, 1, rows, c, F32, DEV)
    dx2 = Act.empty(1, 1, rows, c, dtype, DEV) if dtype != F32 else None
    dg, db = torch.zeros(c, device=DEV), torch.zeros(c, device=DEV)
    ops.layer_norm_bwd(Act(x, 1, 1, rows), Act(dy, 1, 1, rows), gamma, eps, dx, ops.Workspace(DEV),
                       add=Act(add, 1, 1, rows), dx2=dx2, dgamma=dg, dbeta=db)
    ref = xr.grad + add
    assert rel_l2(dx.buf, ref) < 3e-5
    if dx2 is not None:
        assert rel_l2(dx2.buf.float(), ref) < 8e-3
    assert rel_l2(dg, gr.grad) < 1e-4
    assert rel_l2(db, br.grad) < 1e-4


# ---- attention ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype,tc", [(F32, False), (BF16, False), (BF16, True), (BF16, "auto")])
@pytest.mark.parametrize("b,heads,n_q,n_kv", [(2, 8, 256, 256), (3, 16, 64, 64), (2, 8, 1024, 4), (1, 4, 100, 37),
                                               (2, 8, 1024, 1024), (1, 3, 128, 320)])
def test_attention_bwd(dtype, tc, b, heads, n_q, n_kv):
    if tc is True and (n_q % 64 or n_kv % 64):
        pytest.skip("tensor-core backward: self-attention shapes only (multiples of 64)")
    auto = tc == "auto"     # the dispatcher's own choice: register-resident cross-attention kernel for n_kv <= 4
    tc = tc is True
    hd = 32
    C = heads * hd
    scale = hd ** -0.5
    q = torch.randn(b * n_q, C, generator=g(25)).to(DEV).to(dtype)
    kv = torch.randn(b * n_kv, 2 * C, generator=g(26)).to(DEV).to(dtype)
    dout = torch.randn(b * n_q, C, generator=g(27)).to(DEV).to(dtype)

    def heads_view(t, n):
        return t.float().reshape(b, n, heads, hd).permute(0, 2, 1, 3)

    qr = heads_view(q, n_q).requires_grad_(True)
    kr = heads_view(kv[:, :C], n_kv).requires_grad_(True)
    vr = heads_view(kv[:, C:], n_kv).requires_grad_(True)
    o_ref = torch.softmax(qr @ kr.transpose(-1, -2) * scale, dim=-1) @ vr
    o_ref.backward(heads_view(dout, n_q))
    qa = Act(q, b, 1, n_q)
    kva = Act(kv, b, 1, n_kv)
    out = Act.empty(b, 1, n_q, C, dtype, DEV)
    lse = torch.empty(b, heads, n_q, device=DEV) if tc else None
    ops.attention(qa, kva.cols(0, C), kva.cols(C, C), out, batch=b, heads=heads, head_dim=hd, n_q=n_q, n_kv=n_kv,
                  scale=scale, lse=lse)
    if tc:   # the forward's log2-domain log-sum-exp
        ref_lse = torch.logsumexp(qr.detach() @ kr.detach().transpose(-1, -2) * scale, dim=-1) / math.log(2.0)
        assert float((lse - ref_lse).abs().max()) < 2e-2
    dq = Act.empty(b, 1, n_q, C, dtype, DEV)
    dkv = Act.empty(b, 1, n_kv, 2 * C, dtype, DEV)
    ops.attention_bwd(qa, kva.cols(0, C), kva.cols(C, C), out, Act(dout, b, 1, n_q), dq, dkv.cols(0, C),
                      dkv.cols(C, C), ops.Workspace(DEV), batch=b, heads=heads, head_dim=hd, n_q=n_q, n_kv=n_kv,
                      scale=scale, lse=lse, impl=L.IMPL_AUTO if auto else (L.IMPL_TCGEN05 if tc else L.IMPL_SIMT))

    def flat(t, n):
        return t.permute(0, 2, 1, 3).reshape(b * n, C)

    # bf16: `out` is bf16-rounded before D = dO.O; the tensor-core path also rounds P and dS to bf16 (as FA2 does)
    t = 3e-5 if dtype == F32 else (2e-2 if tc else 1.2e-2)
    assert rel_l2(dq.buf.float(), flat(qr.grad, n_q)) < t
    assert rel_l2(dkv.buf[:, :C].float(), flat(kr.grad, n_kv)) < t
    assert rel_l2(dkv.buf[:, C:].float(), flat(vr.grad, n_kv)) < t


# ---- elementwise ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [F32, BF16])
def test_geglu_fwd_bwd(dtype):
    rows, inner = 777, 1024
    pre = torch.randn(rows, 2 * inner, generator=g(28)).to(DEV).to(dtype)
    dout = torch.randn(rows, inner, generator=g(29)).to(DEV).to(dtype)
    pr = pre.float().requires_grad_(True)
    v, gate = pr.chunk(2, dim=-1)
    ref = v * F.gelu(gate)
    ref.backward(dout.float())
    out = ops.geglu(Act(pre, 1, 1, rows), Act.empty(1, 1, rows, inner, dtype, DEV))
    dpre = ops.geglu_bwd(Act(pre, 1, 1, rows), Act(dout, 1, 1, rows), Act.empty(1, 1, rows, 2 * inner, dtype, DEV))
    t = 1e-5 if dtype == F32 else 4e-3
    assert rel_l2(out.buf.float(), ref) < t
    assert rel_l2(dpre.buf.float(), pr.grad) < t


@pytest.mark.parametrize("dtype", [F32, BF16])
def test_silu_fwd_bwd(dtype):
    rows, c = 64, 1024
    x = torch.randn(rows, c, generator=g(30)).to(DEV) * 3
    dy = torch.randn(rows, c, generator=g(31)).to(DEV).to(dtype)
    xr = x.clone().requires_grad_(True)
    F.silu(xr).backward(dy.float())
    y = ops.silu(Act(x, 1, 1, rows), Act.empty(1, 1, rows, c, dtype, DEV))
    dx = ops.silu_bwd(Act(x, 1, 1, rows), Act(dy, 1, 1, rows), Act.empty(1, 1, rows, c, dtype, DEV))
    t = 1e-5 if dtype == F32 else 4e-3
    assert rel_l2(y.buf.float(), F.silu(x)) < t
    assert rel_l2(dx.buf.float(), xr.grad) < t


@pytest.mark.parametrize("dtype", [F32, BF16])
@pytest.mark.parametrize("segs,rps,c", [(1, 131072, 256), (64, 1024, 512), (3, 77, 40), (1, 64, 11008)])
def test_colsum(dtype, segs, rps, c):
    x = torch.randn(segs * rps, c + 8, generator=g(32)).to(DEV).to(dtype)
    xa = Act(x, segs, 1, rps, c, 8)
    base = torch.randn(segs, c + 4, generator=g(33)).to(DEV)
    out = base.clone()
    ops.colsum(xa, out, ops.Workspace(DEV), segs=segs, col0=4, accumulate=True)
    ref = x[:, 8:].float().reshape(segs, rps, c).sum(1)
    assert rel_l2(out[:, 4:] - base[:, 4:], ref) < 2e-5
    assert bool((out[:, :4] == base[:, :4]).all())


def test_cfg_mse_bwd():
    b, per = 6, 4096
    eu = torch.randn(b, 4, 32, 32, generator=g(34)).to(DEV).requires_grad_(True)
    ec = torch.randn(b, 4, 32, 32, generator=g(35)).to(DEV).requires_grad_(True)
    tgt = torch.randn(b, 4, 32, 32, generator=g(36)).to(DEV)
    w = torch.rand(b, generator=g(37)).to(DEV)
    guided = eu + 2.0 * (ec - eu)
    ls = ((guided - tgt) ** 2).mean(dim=[1, 2, 3])
    (ls * w).sum().backward()
    deu, dec = ops.cfg_mse_bwd(ec.detach(), tgt, w, e_uncond=eu.detach(), cfg_scale=2.0)
    assert rel_l2(deu, eu.grad) < 1e-6 and rel_l2(dec, ec.grad) < 1e-6
    ec.grad = None
    (((ec - tgt) ** 2).mean(dim=[1, 2, 3]) * w).sum().backward()
    _, dec = ops.cfg_mse_bwd(ec.detach(), tgt, w)
    assert rel_l2(dec, ec.grad) < 1e-6
