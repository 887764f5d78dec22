"""Per-kernel microbenchmark at the headline workload's shapes (stdiff UNet, UNet batch 128).

  python tools/microbench.py                       # event-timed table (L2 flushed between launches)
  python tools/microbench.py --only gemm_o1_l0 --ncu   # 2 launches between cudaProfilerStart/Stop, for
                                                       # `ncu --profile-from-start off --set full ...`
Reports algorithmic TFLOP/s and algorithmic GB/s (bytes each operand must cross HBM once) per launch.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from ealdm_b200 import _lib as L, ops  # noqa: E402
from ealdm_b200.ops import Act, ConvIn  # noqa: E402

DEV = torch.device("cuda", 0)
BF, F32 = torch.bfloat16, torch.float32


def rnd(rows, c, dt):
    return (torch.randn(rows, c, device=DEV, dtype=F32) * 0.5).to(dt)


def bsz(t):
    return t.numel() * t.element_size()


def make_gemm(M, K, N, *, out_dt=BF, res_dt=None, out2=False, act=L.ACT_NONE, bias=True):
    x = Act(rnd(M, K, BF), 1, 1, M)
    w = rnd(N, K, BF) * 0.1
    n_cols = N // 2 if act == L.ACT_GEGLU else N
    out = Act(torch.empty(M, n_cols, device=DEV, dtype=out_dt), 1, 1, M)
    res = Act(rnd(M, n_cols, res_dt), 1, 1, M) if res_dt is not None else None
    o2 = Act(torch.empty(M, n_cols, device=DEV, dtype=BF), 1, 1, M) if out2 else None
    b = torch.randn(N, device=DEV) if bias else None
    nbytes = bsz(x.buf) + bsz(w) + bsz(out.buf) + (bsz(res.buf) if res else 0) + (bsz(o2.buf) if o2 else 0)

    def run():
        ops.linear(x, w, out, bias=b, residual=res, out2=o2, act=act)
    return run, 2.0 * M * N * K, nbytes


def make_conv3(n, h, w_, cin, cout, *, out_dt=F32, rowvec=True, res_dt=None, skip_c=0):
    x = Act(rnd(n * h * w_, cin, BF), n, h, w_)
    K = 9 * cin + skip_c
    wt = rnd(cout, K, BF) * 0.05
    out = Act(torch.empty(n * h * w_, cout, device=DEV, dtype=out_dt), n, h, w_)
    rv = torch.randn(n, cout, device=DEV) if rowvec else None
    res = Act(rnd(n * h * w_, cout, res_dt), n, h, w_) if res_dt is not None else None
    srcs = [ConvIn(x, 3, 1, 1)]
    nbytes = bsz(x.buf) + bsz(wt) + bsz(out.buf) + (bsz(res.buf) if res else 0)
    if skip_c:
        xs = Act(rnd(n * h * w_, skip_c, BF), n, h, w_)
        srcs.append(ConvIn(xs, 1, 1, 0))
        nbytes += bsz(xs.buf)
    b = torch.randn(cout, device=DEV)

    def run():
        ops.conv(srcs, wt, out, bias=b, rowvec=rv, residual=res)
    return run, 2.0 * n * h * w_ * cout * K, nbytes


def make_conv3_gna(n, h, w_, cin, cout, *, only=True, res=False):
    """3x3 conv whose epilogue applies GroupNorm32 (+ SiLU): only=True writes just the normalised bf16 tensor (a
    ResBlock's conv1 -> GroupNorm -> SiLU), only=False the fp32 result + the normalised copy (conv2 -> transformer norm)."""
    x = Act(rnd(n * h * w_, cin, BF), n, h, w_)
    wt = rnd(cout, 9 * cin, BF) * 0.05
    b = torch.randn(cout, device=DEV)
    gamma, beta = torch.ones(cout, device=DEV), torch.zeros(cout, device=DEV)
    rv = torch.randn(n, cout, device=DEV) if only else None
    r = Act(rnd(n * h * w_, cout, F32), n, h, w_) if res else None
    if only:
        out = Act.empty(n, h, w_, cout, BF, DEV).with_gn_partial()
        o2 = None
    else:
        out = Act.empty(n, h, w_, cout, F32, DEV).with_gn_partial()
        o2 = Act.empty(n, h, w_, cout, BF, DEV)
    nbytes = bsz(x.buf) + bsz(wt) + bsz(out.buf) + (bsz(r.buf) if r else 0) + (bsz(o2.buf) if o2 else 0)

    def run():
        ops.conv([ConvIn(x, 3, 1, 1)], wt, out, bias=b, rowvec=rv, residual=r, out2=o2,
                 gn_apply=(gamma, beta, 1e-5, 32, only, only))
    return run, 2.0 * n * h * w_ * cout * 9 * cin, nbytes


def make_attn(b, heads, n, dh=32):
    C_ = heads * dh
    qkv = Act(rnd(b * n, 3 * C_, BF), b, 1, n)
    o = Act(torch.empty(b * n, C_, device=DEV, dtype=BF), b, 1, n)

    def run():
        ops.attention(qkv.cols(0, C_), qkv.cols(C_, C_), qkv.cols(2 * C_, C_), o, batch=b, heads=heads, head_dim=dh,
                      n_q=n, n_kv=n, scale=dh ** -0.5)
    return run, 4.0 * b * heads * n * n * dh, bsz(qkv.buf) + bsz(o.buf)


def make_xattn(b, heads, n, nk=4, dh=32):
    C_ = heads * dh
    q = Act(rnd(b * n, C_, BF), b, 1, n)
    kv = Act(rnd(b * nk, 2 * C_, BF), b, 1, nk)
    o = Act(torch.empty(b * n, C_, device=DEV, dtype=BF), b, 1, n)

    def run():
        ops.attention(q, kv.cols(0, C_), kv.cols(C_, C_), o, batch=b, heads=heads, head_dim=dh, n_q=n, n_kv=nk,
                      scale=dh ** -0.5)
    return run, 4.0 * b * heads * n * nk * dh, bsz(q.buf) + bsz(o.buf)


def make_gn(n, hw, c, x_dt=F32, silu=True):
    x = Act(rnd(n * hw, c, x_dt), n, 1, hw)
    y = Act(torch.empty(n * hw, c, device=DEV, dtype=BF), n, 1, hw)
    g, b = torch.randn(c, device=DEV), torch.randn(c, device=DEV)
    ws = ops.group_norm_workspace(n, hw, c, DEV)

    def run():
        ops.group_norm(x, g, b, 1e-5, y, ws, silu=silu)
    return run, 0.0, bsz(x.buf) + bsz(y.buf)


def make_gn_partial(n, hw, c, silu=True):
    """GroupNorm fed by the conv epilogue's partial statistics: a single streaming pass."""
    side = int(hw ** 0.5)
    x = Act(rnd(n * hw, c, F32), n, side, side).with_gn_partial()
    ops.gn_partial(x)
    y = Act.empty(n, side, side, c, BF, DEV)
    gamma, beta = torch.ones(c, device=DEV), torch.zeros(c, device=DEV)

    def run():
        ops.group_norm(x, gamma, beta, 1e-5, y, silu=silu)
    return run, 0.0, bsz(x.buf) + bsz(y.buf)


def make_ln(rows, c, x_dt=F32):
    x = Act(rnd(rows, c, x_dt), 1, 1, rows)
    y = Act(torch.empty(rows, c, device=DEV, dtype=BF), 1, 1, rows)
    g, b = torch.randn(c, device=DEV), torch.randn(c, device=DEV)

    def run():
        ops.layer_norm(x, g, b, 1e-5, y)
    return run, 0.0, bsz(x.buf) + bsz(y.buf)


def make_ff_fused(M, out_dt=BF):
    """FF1 -> GEGLU -> FF2 -> + residual in one kernel (c = 256): algorithmic bytes = x + weights + residual + out."""
    x = Act(rnd(M, 256, BF), 1, 1, M)
    w1, b1 = rnd(2048, 256, BF) * 0.1, torch.randn(2048, device=DEV)
    w2, b2 = rnd(256, 1024, BF) * 0.1, torch.randn(256, device=DEV)
    res = Act(rnd(M, 256, F32), 1, 1, M)
    out = Act(torch.empty(M, 256, device=DEV, dtype=out_dt), 1, 1, M)

    def run():
        ops.ff_geglu_fused(x, w1, b1, w2, b2, res, out)
    return run, 2.0 * M * (2048 * 256 + 256 * 1024), bsz(x.buf) + bsz(w1) + bsz(w2) + bsz(res.buf) + bsz(out.buf)


N = 128  # UNet batch of the headline config (B=64, CFG)
CASES = {
    # level 0 transformer GEMMs (131072 tokens x 256)
    "gemm_q2_l0": lambda: make_gemm(N * 1024, 256, 256, bias=False),
    "gemm_qkv_l0": lambda: make_gemm(N * 1024, 256, 768, bias=False),
    "gemm_o1_l0": lambda: make_gemm(N * 1024, 256, 256, out_dt=F32, res_dt=F32),
    "gemm_projout_l0": lambda: make_gemm(N * 1024, 256, 256, out_dt=F32, res_dt=F32, out2=True),
    "gemm_ff1_l0": lambda: make_gemm(N * 1024, 256, 2048, act=L.ACT_GEGLU),
    "gemm_ff2_l0": lambda: make_gemm(N * 1024, 1024, 256, out_dt=F32, res_dt=F32),
    "gemm_ff2bf_l0": lambda: make_gemm(N * 1024, 1024, 256, out_dt=BF, res_dt=F32),
    "ff_fused_l0": lambda: make_ff_fused(N * 1024),
    # level 1 / 2
    "gemm_o1_l1": lambda: make_gemm(N * 256, 512, 512, out_dt=F32, res_dt=F32),
    "gemm_ff1_l1": lambda: make_gemm(N * 256, 512, 4096, act=L.ACT_GEGLU),
    "gemm_o1_l2": lambda: make_gemm(N * 64, 1024, 1024, out_dt=F32, res_dt=F32),
    "gemm_ff1_l2": lambda: make_gemm(N * 64, 1024, 8192, act=L.ACT_GEGLU),
    # ResBlock convs
    "conv3_256_l0": lambda: make_conv3(N, 32, 32, 256, 256),
    "conv3_256_l0_res": lambda: make_conv3(N, 32, 32, 256, 256, rowvec=False, res_dt=F32),
    "conv3_256_l0_gna": lambda: make_conv3_gna(N, 32, 32, 256, 256),
    "conv3_256_l0_res_gna": lambda: make_conv3_gna(N, 32, 32, 256, 256, only=False, res=True),
    "conv3_512_l1_gna": lambda: make_conv3_gna(N, 16, 16, 512, 512),
    "conv3_1024_l2_gna": lambda: make_conv3_gna(N, 8, 8, 1024, 1024),
    "conv3_512_l0": lambda: make_conv3(N, 32, 32, 512, 256),
    "conv3_512_l1": lambda: make_conv3(N, 16, 16, 512, 512),
    "conv3_1024_l2": lambda: make_conv3(N, 8, 8, 1024, 1024),
    "conv3_skip_l0": lambda: make_conv3(N, 32, 32, 256, 256, rowvec=False, skip_c=512),
    # attention
    "attn_l0": lambda: make_attn(N, 8, 1024),
    "attn_l1": lambda: make_attn(N, 16, 256),
    "attn_l2": lambda: make_attn(N, 32, 64),
    "xattn_l0": lambda: make_xattn(N, 8, 1024),
    "xattn_l1": lambda: make_xattn(N, 16, 256),
    "xattn_l2": lambda: make_xattn(N, 32, 64),
    # norms
    "gn_256_l0": lambda: make_gn(N, 1024, 256),
    "gn_256_l0_partial": lambda: make_gn_partial(N, 1024, 256),
    "gn_512_l0_partial": lambda: make_gn_partial(N, 1024, 512),
    "gn_1024_l1_partial": lambda: make_gn_partial(N, 256, 1024),
    "gn_2048_l2_partial": lambda: make_gn_partial(N, 64, 2048),
    "gn_256_l0_bf16in": lambda: make_gn(N, 1024, 256, x_dt=BF),
    "gn_512_l0": lambda: make_gn(N, 1024, 512),
    "gn_768_l0": lambda: make_gn(N, 1024, 768),
    "gn_1024_l1": lambda: make_gn(N, 256, 1024),
    "gn_2048_l2": lambda: make_gn(N, 64, 2048),
    "ln_256_l0": lambda: make_ln(N * 1024, 256),
    "ln_512_l1": lambda: make_ln(N * 256, 512),
    "ln_1024_l2": lambda: make_ln(N * 64, 1024),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None, help="comma-separated case names (prefix match)")
    ap.add_argument("--ncu", action="store_true")
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    torch.cuda.set_device(0)
    names = list(CASES)
    if args.only:
        pref = args.only.split(",")
        names = [n for n in names if any(n.startswith(p) for p in pref)]
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=DEV)
    for name in names:
        run, flops, nbytes = CASES[name]()
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        if args.ncu:
            torch.cuda.cudart().cudaProfilerStart()
            run()
            run()
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaProfilerStop()
            print(f"{name}: profiled")
            continue
        ms = []
        for _ in range(args.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        ms.sort()
        med = ms[len(ms) // 2]
        print(f"{name:18s} {med * 1e3:9.1f} us  min {ms[0] * 1e3:9.1f} us  {flops / (med * 1e-3) / 1e12:8.1f} TF/s  "
              f"{nbytes / (med * 1e-3) / 1e9:8.1f} GB/s  ({nbytes / 1e6:.0f} MB)", flush=True)
        del run
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
