"""One eager UNet forward (stdiff config, bf16, UNet batch 2B) between cudaProfilerStart/Stop, for
`ncu --profile-from-start off ...`.  Also prints an event-timed per-op-class breakdown when run plain."""
import argparse
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from ealdm_b200 import configs as CFG, ops  # noqa: E402
from ealdm_b200.synthetic import init_synthetic_  # noqa: E402
from ealdm_b200.unet import UNetModel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--breakdown", action="store_true")
    ap.add_argument("--plain", action="store_true", help="a plain forward (no guidance pair, context projected inside)")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    unet = UNetModel(**CFG.UNET_STDIFF).to(dev).eval()
    init_synthetic_(unet, 0)
    n = 2 * args.batch
    g = torch.Generator().manual_seed(1)
    x = torch.cat([torch.randn(args.batch, 4, 32, 32, generator=g).to(dev)] * 2)   # as the guided sampler: cat([x] * 2)
    c = torch.randn(n, 4, 512, generator=g).to(dev)
    t = torch.full((n,), 981, device=dev, dtype=torch.long)
    import contextlib
    with contextlib.ExitStack() as stack:
        if not args.plain:   # the forward of one guided DDIM step: shared prefix, context projected by the warm-up
            stack.enter_context(unet.sampling_scope())
            stack.enter_context(unet.cfg_pair())
        for _ in range(2):
            unet(x, t, context=c)
        torch.cuda.synchronize()
        run(args, unet, x, t, c, n)


def run(args, unet, x, t, c, n):
    if args.breakdown:
        rec = defaultdict(list)
        import ealdm_b200.unet as U

        def wrap(name, fn, key=None):
            def w(*a, **k):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                r = fn(*a, **k)
                e1.record()
                rec[name if key is None else key(name, a, k)].append((e0, e1))
                return r
            return w

        def conv_key(name, a, k):
            srcs, weight, out = a[0], a[1], a[2]
            s0 = srcs[0]
            kind = "conv3x3" if s0.ksize == 3 else "gemm"
            if len(srcs) > 1:
                kind += "+skip1x1"
            if s0.stride == 2:
                kind += "_s2"
            tc = s0.x.dtype == torch.bfloat16 and s0.x.c % 64 == 0
            return f"{kind}[{'tc' if tc else 'simt'}] M={out.rows} N={weight.shape[0]} K={weight.shape[1]}"

        orig = {k: getattr(ops, k) for k in ("conv", "group_norm", "layer_norm", "attention", "upsample_nearest2x",
                                              "copy2d", "nchw_to_nhwc", "nhwc_to_nchw", "timestep_embedding")}
        ops.conv = wrap("conv", orig["conv"], conv_key)
        for k in orig:
            if k != "conv":
                setattr(ops, k, wrap(k, orig[k], (lambda name, a, kw: f"attention n_kv={kw['n_kv']} n_q={kw['n_q']}")
                                     if k == "attention" else None))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        unet(x, t, context=c)
        e1.record()
        torch.cuda.synchronize()
        total = e0.elapsed_time(e1)
        rows = sorted(((sum(a.elapsed_time(b) for a, b in v), len(v), k) for k, v in rec.items()), reverse=True)
        print(f"forward {total:.2f} ms at UNet batch {n}")
        acc = 0.0
        for ms, cnt, k in rows:
            acc += ms
            fl = ""
            if "M=" in k:
                M, N, K = (int(p.split("=")[1]) for p in k.split()[1:4])
                fl = f"  {2.0 * M * N * K * cnt / (ms * 1e-3) / 1e12:7.1f} TF/s"
            print(f"{ms:8.3f} ms  x{cnt:<3d} {100 * ms / total:5.1f}%  {k}{fl}")
        print(f"sum of ops {acc:.2f} ms")
        for k, v in orig.items():
            setattr(ops, k, v)
        return
    torch.cuda.profiler.start()
    unet(x, t, context=c)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
