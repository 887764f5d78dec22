"""Event-timed per-operator breakdown of one training step (stdiff UNet, bf16, batch 32 -> UNet batch 64):
forward with saved activations + hand-written backward.  Also times engine construction (weight packing)."""
import argparse
import os
import sys
import time
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from ealdm_b200 import configs as CFG, ops  # noqa: E402
from ealdm_b200.ddpm import LatentDiffusion  # noqa: E402
from ealdm_b200.synthetic import init_synthetic_  # noqa: E402
from ealdm_b200.train import UNetTrainEngine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--ncu", action="store_true", help="cudaProfilerStart/Stop around one step, no breakdown")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    ld = LatentDiffusion(unet_config={"target": "ealdm_b200.unet.UNetModel", "params": dict(CFG.UNET_STDIFF)},
                         cond_stage_config={"target": "torch.nn.Identity"}, conditioning_key="crossattn",
                         **CFG.DIFFUSION).to(dev).train()
    unet = ld.model.diffusion_model
    init_synthetic_(unet, 0)
    B = args.batch
    g = torch.Generator().manual_seed(1)
    x0 = torch.randn(B, 4, 32, 32, generator=g).to(dev)
    c2 = torch.randn(2 * B, 4, 512, generator=g).to(dev)
    t = torch.randint(0, 1000, (B,), generator=g).to(dev)
    noise = torch.randn(B, 4, 32, 32, generator=g).to(dev)

    def step():
        for p in unet.parameters():
            if p.grad is not None:
                p.grad.zero_()
        loss, _ = ld.p_losses(x0, c2, t, noise=noise)
        loss.backward()
        return loss

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    if args.ncu:
        torch.cuda.profiler.start()
        step()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    t0 = time.perf_counter()
    UNetTrainEngine(unet, torch.bfloat16)
    torch.cuda.synchronize()
    print(f"engine construction (weight packing for forward + dgrad): {1e3 * (time.perf_counter() - t0):.1f} ms wall")

    rec = defaultdict(list)
    names = ["conv", "conv_wgrad", "group_norm", "group_norm_bwd", "layer_norm", "layer_norm_bwd", "attention",
             "attention_bwd", "geglu", "geglu_bwd", "silu", "silu_bwd", "colsum", "zero_insert2x", "sumpool2x2",
             "upsample_nearest2x", "copy2d", "nchw_to_nhwc", "nhwc_to_nchw", "timestep_embedding", "cfg_mse",
             "cfg_mse_bwd", "q_sample"]
    orig = {k: getattr(ops, k) for k in names}

    def key(name, a, k):
        if name == "conv":
            srcs, weight, out = a[0], a[1], a[2]
            kind = "conv3x3" if srcs[0].ksize == 3 else "gemm"
            return f"{kind} M={out.rows} N={weight.shape[0]} K={weight.shape[1]}", 2.0 * out.rows * weight.shape[0] * weight.shape[1]
        if name == "conv_wgrad":
            x, dy = a[0], a[1]
            ks = k.get("ksize", 1)
            return f"wgrad{ks}x{ks} M={x.n * dy.h * dy.w} N={dy.c} K={ks * ks * x.c}", 2.0 * x.n * dy.h * dy.w * dy.c * ks * ks * x.c
        if name in ("attention", "attention_bwd"):
            return f"{name} n_q={k['n_q']} n_kv={k['n_kv']}", 0.0
        return name, 0.0

    def wrap(name, fn):
        def w(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            nested = getattr(wrap, "depth", 0)
            wrap.depth = nested + 1
            if nested == 0:
                e0.record()
            r = fn(*a, **k)
            if nested == 0:
                e1.record()
                kk, fl = key(name, a, k)
                rec[kk].append((e0, e1, fl))
            wrap.depth = nested
            return r
        return w

    for k_ in names:
        setattr(ops, k_, wrap(k_, orig[k_]))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw = time.perf_counter()
    e0.record()
    step()
    e1.record()
    torch.cuda.synchronize()
    wall = 1e3 * (time.perf_counter() - tw)
    for k_, v in orig.items():
        setattr(ops, k_, v)
    total = e0.elapsed_time(e1)
    print(f"training step (fwd+bwd, no optimizer) {total:.2f} ms GPU, {wall:.1f} ms wall, UNet batch {2 * B}")
    groups = defaultdict(lambda: [0.0, 0, 0.0])
    for k_, v in rec.items():
        cls = k_.split()[0]
        ms = sum(a.elapsed_time(b) for a, b, _ in v)
        groups[cls][0] += ms
        groups[cls][1] += len(v)
        groups[cls][2] += sum(f for _, _, f in v)
    print("--- by operator class ---")
    for cls, (ms, cnt, fl) in sorted(groups.items(), key=lambda kv: -kv[1][0]):
        tf = f"  {fl / (ms * 1e-3) / 1e12:7.1f} TF/s" if fl > 0 else ""
        print(f"{ms:8.3f} ms  x{cnt:<4d} {100 * ms / total:5.1f}%  {cls}{tf}")
    print("--- top individual shapes ---")
    rows = sorted(((sum(a.elapsed_time(b) for a, b, _ in v), len(v), k_, sum(f for _, _, f in v)) for k_, v in rec.items()),
                  reverse=True)
    for ms, cnt, k_, fl in rows[:40]:
        tf = f"  {fl / (ms * 1e-3) / 1e12:7.1f} TF/s" if fl > 0 else ""
        print(f"{ms:8.3f} ms  x{cnt:<3d} {100 * ms / total:5.1f}%  {k_}{tf}")


if __name__ == "__main__":
    main()
