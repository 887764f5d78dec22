"""Parity of the tcgen05 flash attention against torch (fp32 math on the bf16-rounded inputs)."""
import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ealdm_b200 import _lib as L, ops
from ealdm_b200.ops import Act

def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())

def run(b, heads, n, legacy=False):
    hd, C = 32, heads * 32
    g = torch.Generator().manual_seed(b * 1000 + n)
    qkv = torch.randn(b * n, 3 * C, generator=g).cuda().bfloat16()
    A = Act(qkv, b, 1, n)
    out = Act.empty(b, 1, n, C, torch.bfloat16, "cuda")
    lse = torch.empty(b, heads, n, device="cuda")
    if legacy:   # per-head [q;k;v] channel interleave (QKVAttentionLegacy): head stride 96
        span = 3 * C - 2 * hd
        q, k, v = A.cols(0, span), A.cols(hd, span), A.cols(2 * hd, span)
        hs = 3 * hd
        x = qkv.float().reshape(b, n, heads, 3, hd)
        qr, kr, vr = (x[:, :, :, i].permute(0, 2, 1, 3) for i in range(3))
    else:
        q, k, v = A.cols(0, C), A.cols(C, C), A.cols(2 * C, C)
        hs = hd
        x = qkv.float().reshape(b, n, 3, heads, hd)
        qr, kr, vr = (x[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    ops.attention(q, k, v, out, batch=b, heads=heads, head_dim=hd, n_q=n, n_kv=n, scale=hd ** -0.5,
                  head_stride_q=hs, head_stride_kv=hs, impl=L.IMPL_TCGEN05, lse=lse)
    torch.cuda.synchronize()
    sc = qr @ kr.transpose(-1, -2) * hd ** -0.5
    ref = (torch.softmax(sc, -1) @ vr).permute(0, 2, 1, 3).reshape(b * n, C)
    ref_lse = torch.logsumexp(sc, -1) / math.log(2.0)
    print(f"b={b} h={heads} n={n} legacy={legacy}: out rel_l2 {rel(out.buf.float(), ref):.3e}  lse max abs {float((lse - ref_lse).abs().max()):.3e}")

for args in [(1, 1, 128), (2, 8, 128), (2, 8, 1024), (3, 16, 256), (5, 4, 384)]:
    run(*args)
run(2, 8, 256, legacy=True)
