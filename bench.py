#!/usr/bin/env python
"""Headline benchmark: latent samples/sec of 50-step conditioned DDIM sampling (32x32x4 latent, 256 px,
classifier-free guidance 2.0) on the stdiff_cin-ldm-vq-f8 environment-conditioned UNet, batch 64 per GPU
(BASELINE.json configs[1]); synthetic latents / conditioning, random-init weights.

  python bench.py --gpus N --steps K --warmup W           one rank per GPU under torchrun for N > 1
  python bench.py --impl reference ...                    the reference's CPU implementation (oracle port)

A "step" is ONE full DDIMSampler.sample() call over one batch (50 UNet evaluations at UNet batch 2B
plus 50 fused DDIM updates) -- the timed region the reference's own throughput print uses
(scripts/sample_diffusion.py:90-104, first-stage decode excluded).
Prints one JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "latent samples/sec (50-step DDIM, 256px)"
UNIT = "samples/s"
WORKLOAD = ("stdiff_cin-ldm-vq-f8 UNet (394.98 M params, cross-attention on [B,4,512] STDiff conditioning), 50-step DDIM, "
            "CFG 2.0, eta 0, 32x32x4 latent, batch 64 per GPU (BASELINE.json configs[1])")


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_sustained": float(p.get("bf16_tflops_sustained", p.get("bf16_tflops"))),
                "bf16_burst": float(p["bf16_tflops"]), "hbm": float(p["hbm_gbs"]), "source": "measured"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        load = [v for v in sm if v > 0.5 * max(sm)] if sm else []
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The reference is Python and
    cannot travel to the GPU box, so this times the oracle port (a torch-fp32 restatement pinned against
    the reference's outputs) on all host cores.  Each step is a bounded sample of the workload: ONE DDIM
    step (UNet at batch 2B with CFG + update) at B = 2, extrapolated linearly to 50 steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch

    from ealdm_b200 import configs as CFG
    from oracle import diffusion as OD
    from oracle import unet as OU

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B, S = 2, 50
    sd = OU.synthetic_state_dict(OU.unet_param_shapes(CFG.UNET_STDIFF), seed=2)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, 4, 32, 32, generator=g)
    c, uc = torch.randn(B, 4, 512, generator=g), torch.randn(B, 4, 512, generator=g)
    buf = OD.register_schedule(1000, CFG.DIFFUSION["linear_start"], CFG.DIFFUSION["linear_end"])
    sched = OD.make_ddim_schedule(buf["alphas_cumprod"], S, 0.0)
    apply_model = lambda xx, tt, cc: OU.unet_forward(sd, CFG.UNET_STDIFF, xx, tt, cc)  # noqa: E731

    def one_step(i):
        index = S - 1 - (i % S)
        t = torch.full((B,), int(sched["ddim_timesteps"][index]), dtype=torch.long)
        with torch.no_grad():
            return OD.p_sample_ddim(apply_model, x, c, t, sched, index, torch.zeros_like(x), ugs=2.0, uc=uc)[0]

    for i in range(args.warmup):
        one_step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        one_step(i)
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    value = B / (dt * S)
    sample = (f"bounded sample of the workload: B={B} (not 64), ONE of the {S} DDIM steps per timed step (CFG 2.0, UNet batch "
              f"{2 * B}), fp32, oracle port on {cores} host threads; value = B / (step time x {S}), i.e. linearly "
              f"extrapolated in the step count; per-sample cost at B=64 on a CPU is the same or lower")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "timed_batch": B, "timed_ddim_steps_per_step": 1, "ddim_steps": S,
                       "guidance_scale": 2.0, "same_config": False,
                       "note": "the full workload (B=64, 50 steps) is ~27 min of CPU time per step; see cpu_baseline.sample"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


def cpu_baseline_leg(budget_s=25.0):
    """Oracle port on the host cores, bounded: one DDIM step (CFG) at B=2, extrapolated to 50 steps."""
    import torch

    from ealdm_b200 import configs as CFG
    from oracle import diffusion as OD
    from oracle import unet as OU
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B, S = 2, 50
    sd = OU.synthetic_state_dict(OU.unet_param_shapes(CFG.UNET_STDIFF), seed=2)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, 4, 32, 32, generator=g)
    c, uc = torch.randn(B, 4, 512, generator=g), torch.randn(B, 4, 512, generator=g)
    buf = OD.register_schedule(1000, CFG.DIFFUSION["linear_start"], CFG.DIFFUSION["linear_end"])
    sched = OD.make_ddim_schedule(buf["alphas_cumprod"], S, 0.0)
    apply_model = lambda xx, tt, cc: OU.unet_forward(sd, CFG.UNET_STDIFF, xx, tt, cc)  # noqa: E731
    t = torch.full((B,), 981, dtype=torch.long)
    times = []
    t_start = time.perf_counter()
    with torch.no_grad():
        OD.p_sample_ddim(apply_model, x, c, t, sched, S - 1, torch.zeros_like(x), ugs=2.0, uc=uc)  # warm-up
        while len(times) < 3 and time.perf_counter() - t_start < budget_s:
            t0 = time.perf_counter()
            OD.p_sample_ddim(apply_model, x, c, t, sched, S - 1, torch.zeros_like(x), ugs=2.0, uc=uc)
            times.append(time.perf_counter() - t0)
    dt = min(times)
    return {"value": B / (dt * S), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"oracle port (torch fp32 CPU), B={B}, 1 of {S} DDIM steps with CFG (UNet batch {2 * B}), "
                      f"best of {len(times)}, extrapolated linearly to {S} steps"}


def torch_gpu_baseline_leg(dev, B, S):
    """The GPU-side comparison bar of BASELINE.md section 4: eager PyTorch (cuDNN / cuBLAS kernels, the arithmetic
    provider of the reference) on the SAME B200 -- the oracle port's functional restatement of the reference UNet run
    on cuda for ONE DDIM step's UNet evaluation at the benchmark's UNet batch (2B with CFG), in fp32 with TF32 off,
    fp32 with TF32 on, and under torch.autocast(bf16).  Outside every timed region; reported like cpu_baseline
    (`value` = B / (50 x forward time): the DDIM update is not included, which favours the baseline)."""
    import torch

    from ealdm_b200 import configs as CFG
    from oracle import unet as OU   # checker code, used here as the eager-PyTorch baseline only

    sd = {k: v.to(dev) for k, v in OU.synthetic_state_dict(OU.unet_param_shapes(CFG.UNET_STDIFF), seed=2).items()}
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2 * B, 4, 32, 32, generator=g).to(dev)
    c = torch.randn(2 * B, 4, 512, generator=g).to(dev)
    t = torch.full((2 * B,), 981, dtype=torch.long, device=dev)
    out = {"unit": UNIT, "unet_batch": 2 * B, "kind": "eager PyTorch (cuDNN/cuBLAS) on this GPU, oracle port of the "
           "reference UNet, one UNet evaluation timed and scaled by the 50 DDIM steps"}
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)

    def timed(ctx):
        with torch.no_grad(), ctx:
            OU.unet_forward(sd, CFG.UNET_STDIFF, x, t, c)          # warm-up (cuDNN heuristics, allocator)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(2):
                OU.unet_forward(sd, CFG.UNET_STDIFF, x, t, c)
            e1.record()
            torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 2

    import contextlib
    try:
        for name, tf32, ctx in (("fp32", False, contextlib.nullcontext()), ("tf32", True, contextlib.nullcontext()),
                                ("autocast_bf16", True, torch.autocast("cuda", dtype=torch.bfloat16))):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            ms = timed(ctx)
            out[name] = {"unet_forward_ms": ms, "value": B / (ms * 1e-3 * S)}
    except Exception as ex:   # a baseline must never break the product line
        out["error"] = f"{type(ex).__name__}: {ex}"
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
        del sd
        torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------------
def run_train(args):
    """--workload train: BASELINE.json configs[4] -- one optimisation step of LatentDiffusion on the stdiff UNet:
    q_sample, UNet forward at 2B (classifier-free guidance inside the loss, ddpm.py:1040-1044), eps loss, the
    hand-written backward, the bucketed NCCL gradient all-reduce overlapped with it, AdamW (the reference's
    optimizer, ddpm.py:1409-1431).  Batch 32 per GPU, bf16 operands / fp32 accumulation and master weights."""
    import torch
    import torch.distributed as dist

    from ealdm_b200 import configs as CFG, ops
    from ealdm_b200.ddpm import LatentDiffusion
    from ealdm_b200.synthetic import init_synthetic_
    from ealdm_b200.train import FusedTrainStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    own_group = world > 1 and not dist.is_initialized()
    if own_group:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch if args.batch != 64 else 32
    ld = LatentDiffusion(unet_config={"target": "ealdm_b200.unet.UNetModel", "params": dict(CFG.UNET_STDIFF)},
                         cond_stage_config={"target": "torch.nn.Identity"}, conditioning_key="crossattn",
                         **CFG.DIFFUSION).to(dev).train()
    unet = ld.model.diffusion_model
    init_synthetic_(unet, seed=0)
    unet.set_compute_dtype("bf16")
    fused = FusedTrainStep(ld, use_graph=not args.no_graph, bucket_mb=64.0)
    gb = fused.buckets
    lr = 1.0e-6 * world * B      # main.py:741-745: accumulate * ngpu * bs * base_lr
    if args.torch_optim:         # torch.optim.AdamW(fused=True), no EMA
        opt = torch.optim.AdamW(unet.parameters(), lr=lr, fused=True)
    else:                        # AdamW + EMA shadow (use_ema: True in the reference config) + bf16 weights, one kernel
        from ealdm_b200.optim import FusedAdamWEMA
        opt = FusedAdamWEMA(gb, lr=lr)
    g = torch.Generator().manual_seed(100 + rank)
    x0_h = torch.randn(B, 4, 32, 32, generator=g).pin_memory()
    c2_h = torch.randn(2 * B, 4, 512, generator=g).pin_memory()
    t = torch.randint(0, 1000, (B,), generator=g).to(dev)
    noise = torch.randn(B, 4, 32, 32, generator=g).to(dev)
    phases = {"fwd(+bwd if fused)": [], "bwd/allreduce": [], "opt": []}

    def step(record=False):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        x0 = x0_h.to(dev, non_blocking=True)
        c2 = c2_h.to(dev, non_blocking=True)
        gb.zero_()
        ev[0].record()
        if args.autograd:           # the drop-in path: p_losses under torch.autograd
            loss, _ = ld.p_losses(x0, c2, t, noise=noise)
            ev[1].record()
            loss.backward()
        else:                       # the fused step (one CUDA graph unless --no-graph)
            loss = fused(x0, c2, t, noise)
            ev[1].record()
        if args.torch_optim:
            gb.finish()
            ev[2].record()
            opt.step()
        else:
            gb.finish(average=False)
            ev[2].record()
            opt.step(grads_are_sums=True)
        ev[3].record()
        lv = float(loss)            # device -> host read of the step's result
        if record:
            torch.cuda.synchronize()
            for k, (a, b) in zip(phases.keys(), ((0, 1), (1, 2), (2, 3))):
                phases[k].append(ev[a].elapsed_time(ev[b]))
        return lv

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    l0 = ops.launch_count()
    for _ in range(max(args.warmup, 3)):
        step()
    launches = (ops.launch_count() - l0) // max(args.warmup, 3)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if args.ncu_window:
        torch.cuda.profiler.start()
    e0.record()
    for _ in range(args.steps):
        lv = step()
    e1.record()
    barrier()
    if args.ncu_window:
        torch.cuda.profiler.stop()
    ms = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    clk = clocks.stop() if rank == 0 else None
    for _ in range(2):
        step(record=True)
    ms_per_step = ms / args.steps
    peaks = load_peaks()
    model_tf = 3 * 2 * B * CFG.UNET_STDIFF_GFLOP_PER_SAMPLE / 1e3     # fwd + dgrad + wgrad, per GPU per step
    if rank == 0:
        line = {"metric": "training samples/sec (UNet fwd+bwd, eps loss, AdamW + EMA)", "value": B * world / (ms_per_step * 1e-3),
                "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": "stdiff_cin-ldm-vq-f8 training step, batch 32 per GPU (UNet batch 64), CFG inside "
                                       "the loss, AdamW, bucketed NCCL gradient all-reduce (BASELINE.json configs[4])",
                           "batch_per_gpu": B, "global_batch": B * world, "lr": lr,
                           "step": "autograd (p_losses + loss.backward())" if args.autograd else
                                   ("fused step, eager launches" if args.no_graph else
                                    ("fused step, one CUDA graph" if world == 1 else
                                     "fused step, 5 CUDA-graph segments with the all-reduce launched between them")),
                           "optimizer": "torch.optim.AdamW(fused=True)" if args.torch_optim else
                                        "ealdm_adamw_ema_step (AdamW + EMA + bf16 weights in one pass)",
                           "parallelism": f"data-parallel x{world}, 64 MiB gradient buckets"},
                "clocks": clk, "loss": lv, "gpu_launches": int(launches * args.steps), "launches_per_step": int(launches),
                "phases_ms": {k: sum(v) / len(v) for k, v in phases.items()},
                "model_tflops_per_gpu": model_tf / (ms_per_step * 1e-3),
                "frac_of_sustained_peak": model_tf / (ms_per_step * 1e-3) / peaks["bf16_sustained"],
                "e2e": {"value": B * world / (ms_per_step * 1e-3), "unit": "samples/s",
                        "h2d_bytes_per_step": (x0_h.numel() + c2_h.numel()) * 4, "d2h_bytes_per_step": 4,
                        "note": "the timed step already copies its batch from pinned host memory and reads the loss back"}}
        if not getattr(args, "quiet", False):
            print(json.dumps(line), flush=True)
    else:
        line = None
    unet.grad_ready_hook = None
    if own_group:
        dist.barrier()
        dist.destroy_process_group()
    return line


def run_autoencoder(args):
    """--workload autoencoder: BASELINE.json configs[2] -- AutoencoderKL (f=8, 32x32x4) encode + decode of 256x256
    images, batch 64, bf16 operands.  Algorithmic work: 272.722 GF (encode) + 622.187 GF (decode) per image."""
    import torch

    from ealdm_b200 import configs as CFG, ops
    from ealdm_b200.autoencoder import AutoencoderKL
    from ealdm_b200.synthetic import init_synthetic_

    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    B = args.batch
    ae = AutoencoderKL(ddconfig=dict(CFG.AE_KL_F8_DDCONFIG), embed_dim=4).to(dev).eval()
    init_synthetic_(ae, seed=0)
    ae.set_compute_dtype("bf16")
    g = torch.Generator().manual_seed(1)
    img_h = (torch.rand(B, 3, 256, 256, generator=g) * 2 - 1).pin_memory()
    out_h = torch.empty(B, 3, 256, 256).pin_memory()
    img_d = img_h.to(dev)

    def step(host):
        x = img_h.to(dev, non_blocking=True) if host else img_d
        with torch.no_grad():
            z = ae.encode(x).mode()
            rec = ae.decode(z)
        if host:
            out_h.copy_(rec, non_blocking=True)
            torch.cuda.current_stream().synchronize()    # the caller consumes the host images of every step
        return rec

    res = {}
    l0 = ops.launch_count()
    for host in (False, True):
        for _ in range(max(args.warmup, 3)):
            step(host)
        torch.cuda.synchronize()
        if not host:
            launches = (ops.launch_count() - l0) // max(args.warmup, 3)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step(host)
        e1.record()
        torch.cuda.synchronize()
        res[host] = e0.elapsed_time(e1) / args.steps
    peaks = load_peaks()
    tf = B * (272.722 + 622.187) / 1e3
    line = {"metric": "images/sec (AutoencoderKL encode+decode, 256px)", "value": B / (res[False] * 1e-3),
            "unit": "images/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": res[False],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "autoencoder_kl_32x32x4 (f=8) encode + decode 256x256, batch 64 (BASELINE.json configs[2])",
                       "batch_per_gpu": B},
            "e2e": {"value": B / (res[True] * 1e-3), "unit": "images/s", "h2d_bytes_per_step": img_h.numel() * 4,
                    "d2h_bytes_per_step": out_h.numel() * 4, "ms_per_step": res[True]},
            "gpu_launches": int(launches * args.steps), "tflops_per_gpu": tf / (res[False] * 1e-3),
            "frac_of_sustained_peak": tf / (res[False] * 1e-3) / peaks["bf16_sustained"]}
    if not getattr(args, "quiet", False):
        print(json.dumps(line), flush=True)
    return line


def run_config1(args):
    """--workload config1: BASELINE.json configs[0], the reference's own CPU-runnable case, run EXACTLY (no
    extrapolation) on both sides: uncond_cin-ldm-vq-f8 UNet (AttentionBlock at every level), 10-step DDIM, batch 4,
    32x32x4 latent, eta 0.  GPU: the module API in fp32 parity mode (SIMT kernels) and in bf16 (tcgen05);
    CPU: the oracle port in fp32 on all host cores, the full 10 steps."""
    import torch

    from ealdm_b200 import configs as CFG, ops
    from ealdm_b200.ddim import DDIMSampler
    from ealdm_b200.ddpm import LatentDiffusion
    from oracle import diffusion as OD   # cpu_baseline leg only
    from oracle import unet as OU

    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    B, S = 4, 10
    ld = LatentDiffusion(unet_config={"target": "ealdm_b200.unet.UNetModel", "params": dict(CFG.UNET_UNCOND)},
                         **CFG.DIFFUSION)
    sd = OU.synthetic_state_dict(OU.unet_param_shapes(CFG.UNET_UNCOND), seed=1)
    ld.model.diffusion_model.load_state_dict(sd, strict=True)
    ld = ld.to(dev).eval()
    unet = ld.model.diffusion_model
    sampler = DDIMSampler(ld)
    x_T_h = torch.randn(B, 4, 32, 32, generator=torch.Generator().manual_seed(1)).pin_memory()
    out_h = torch.empty(B, 4, 32, 32).pin_memory()
    res = {}
    l0 = ops.launch_count()
    for mode in ("fp32", "bf16"):
        unet.set_compute_dtype(mode)
        unet.enable_cuda_graph(True)

        def step():
            z, _ = sampler.sample(S=S, batch_size=B, shape=(4, 32, 32), eta=0.0, x_T=x_T_h.to(dev, non_blocking=True),
                                  verbose=False)
            out_h.copy_(z, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return z

        for _ in range(max(args.warmup, 3)):
            z = step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            z = step()
        e1.record()
        torch.cuda.synchronize()
        res[mode] = (e0.elapsed_time(e1) / args.steps, z.cpu())
    launches = ops.launch_count() - l0
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    buf = OD.register_schedule(1000, CFG.DIFFUSION["linear_start"], CFG.DIFFUSION["linear_end"])
    apply_model = lambda xx, tt, cc: OU.unet_forward(sd, CFG.UNET_UNCOND, xx, tt)  # noqa: E731
    t0 = time.perf_counter()
    with torch.no_grad():
        ref, _ = OD.ddim_sample(apply_model, buf["alphas_cumprod"], S, x_T_h.clone(), eta=0.0)
    cpu_s = time.perf_counter() - t0
    rel = lambda a: float((a.double() - ref.double()).norm() / ref.double().norm())  # noqa: E731
    line = {"metric": "latent samples/sec (10-step DDIM, config 1)", "value": B / (res["fp32"][0] * 1e-3), "unit": "samples/s",
            "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": res["fp32"][0],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "uncond_cin-ldm-vq-f8 UNet, 10-step DDIM, batch 4, 32x32x4 latent, eta 0, random-init "
                                   "weights, fp32 parity mode (BASELINE.json configs[0]); host x_T in, host latents out"},
            "bf16": {"value": B / (res["bf16"][0] * 1e-3), "unit": "samples/s", "ms_per_step": res["bf16"][0],
                     "rel_l2_vs_cpu_oracle": rel(res["bf16"][1])},
            "rel_l2_vs_cpu_oracle": rel(res["fp32"][1]), "gpu_launches": int(launches),
            "e2e": {"value": B / (res["fp32"][0] * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": x_T_h.numel() * 4,
                    "d2h_bytes_per_step": out_h.numel() * 4},
            "cpu_baseline": {"value": B / cpu_s, "unit": "samples/s", "cores": cores, "kind": "port",
                             "sample": "the complete workload: oracle port (torch fp32 CPU), B=4, all 10 DDIM steps"}}
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="sample", choices=["sample", "train", "autoencoder", "config1"],
                    help="sample: the headline metric (default); train: BASELINE.json configs[4]; "
                         "autoencoder: BASELINE.json configs[2]")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="latent samples per GPU per step")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="BASELINE.json configs[3]: a FIXED global batch (512) sharded over the ranks -> strong scaling")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the torch-GPU baseline and the configs[2] / configs[4] sub-results of the default line")
    ap.add_argument("--ddim-steps", type=int, default=50)
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--autograd", action="store_true", help="train workload: p_losses under torch.autograd")
    ap.add_argument("--torch-optim", action="store_true", help="train workload: torch.optim.AdamW(fused=True), no EMA")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-decode", action="store_true", help="skip the decode-inclusive images/s figure")
    ap.add_argument("--ncu-window", action="store_true",
                    help="cudaProfilerStart/Stop around the timed region (for `ncu --profile-from-start off`)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "train":
        run_train(args)
        return 0
    if args.workload == "autoencoder":
        run_autoencoder(args)
        return 0
    if args.workload == "config1":
        return run_config1(args)

    import torch
    import torch.distributed as dist

    from ealdm_b200 import configs as CFG, ops
    from ealdm_b200.ddim import DDIMSampler
    from ealdm_b200.ddpm import LatentDiffusion
    from ealdm_b200.parallel import gather_batch, shard
    from ealdm_b200.synthetic import init_synthetic_

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback; use --impl reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # the CPU baseline runs at N = 1 only, BEFORE any GPU work or process group exists: nothing else competes for the
    # host cores (round 1 ran it on rank 0 while the other ranks polled a barrier, which corrupted the N > 1 lines)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_baseline_leg()
        except Exception as ex:  # the oracle is a checker; never let it break the GPU line
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    B, S, ugs = args.batch, args.ddim_steps, 2.0
    strong = args.global_batch > 0
    if strong:
        assert args.global_batch % world == 0, "--global-batch must divide by the number of ranks"
        B = args.global_batch // world
    ld = LatentDiffusion(unet_config={"target": "ealdm_b200.unet.UNetModel", "params": dict(CFG.UNET_STDIFF)},
                         cond_stage_config={"target": "torch.nn.Identity"}, conditioning_key="crossattn",
                         **CFG.DIFFUSION).to(dev).eval()
    unet = ld.model.diffusion_model
    init_synthetic_(unet, seed=0)
    unet.set_compute_dtype("bf16")
    unet.enable_cuda_graph(not args.no_graph)
    sampler = DDIMSampler(ld)

    # global synthetic inputs drawn identically on every rank, then sliced (bit-identical to 1 GPU)
    g = torch.Generator().manual_seed(1)
    Bg = B * world
    x_T_h = torch.randn(Bg, 4, 32, 32, generator=g).pin_memory()
    cond_h = torch.randn(Bg, 4, 512, generator=g).pin_memory()
    uc_h = torch.randn(Bg, 4, 512, generator=g).pin_memory()
    x_T_h, cond_h, uc_h = (shard(t, rank, world).pin_memory() for t in (x_T_h, cond_h, uc_h))
    out_h = torch.empty(B, 4, 32, 32).pin_memory()

    def sample(x_T, cond, uc):
        z, _ = sampler.sample(S=S, batch_size=B, shape=(4, 32, 32), conditioning=cond, eta=0.0, x_T=x_T,
                              verbose=False, unconditional_guidance_scale=ugs, unconditional_conditioning=uc)
        return z

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(n):
            fn()
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # --- device-resident arm ("value") -------------------------------------------------------------
    x_T_d, cond_d, uc_d = x_T_h.to(dev), cond_h.to(dev), uc_h.to(dev)

    def step_resident():
        z = sample(x_T_d, cond_d, uc_d)
        if world > 1:
            gather_batch(z, Bg)      # the single NCCL gather of the job
        return z

    launches0 = ops.launch_count()
    for _ in range(max(args.warmup, 3)):
        step_resident()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    if args.ncu_window:
        torch.cuda.cudart().cudaProfilerStart()
    ms = timed(step_resident, args.steps)
    if args.ncu_window:
        torch.cuda.cudart().cudaProfilerStop()
    clk = clocks.stop() if rank == 0 else None
    ms_per_step = ms / args.steps
    value = Bg / (ms_per_step * 1e-3)

    # --- end-to-end arm ("e2e"): host buffers in, host latents out, copies inside the timed region ----
    def step_e2e():
        xd = x_T_h.to(dev, non_blocking=True)
        cd = cond_h.to(dev, non_blocking=True)
        ud = uc_h.to(dev, non_blocking=True)
        z = sample(xd, cd, ud)
        out_h.copy_(z, non_blocking=True)
        if world > 1:
            gather_batch(z, Bg)
        torch.cuda.current_stream().synchronize()        # the caller consumes the host latents of every step
        return z

    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps) / args.steps
    e2e_value = Bg / (ms_e2e * 1e-3)
    h2d = (x_T_h.numel() + cond_h.numel() + uc_h.numel()) * 4
    d2h = out_h.numel() * 4

    # --- decode-inclusive images/s (reported separately from the headline) -------
    # AutoencoderKL (f=8, 32x32x4 -> 256x256x3) decode of the step's latents, device resident, after sampling
    with_decode = None
    if not args.no_decode:
        from ealdm_b200.autoencoder import AutoencoderKL
        ae = AutoencoderKL(ddconfig=dict(CFG.AE_KL_F8_DDCONFIG), embed_dim=CFG.AE_KL_F8_EMBED_DIM).to(dev).eval()
        init_synthetic_(ae, seed=3)
        ae.set_compute_dtype("bf16")
        zlat = step_resident()
        for _ in range(2):
            ae.decode(zlat)
        ms_dec = timed(lambda: ae.decode(zlat), 3) / 3
        with_decode = {"value": Bg / ((ms_per_step + ms_dec) * 1e-3), "unit": "images/s", "decode_ms": ms_dec,
                       "first_stage": "AutoencoderKL f=8 (autoencoder_kl_32x32x4), bf16, 256x256x3 output"}
        del ae

    # --- roofline of the dominant kernel: the tcgen05 implicit-GEMM conv/linear ----------------------
    # one eager, instrumented UNet forward at the benchmark's UNet batch: CUDA events around every
    # ealdm_conv launch on the launching stream; algorithmic FLOPs = 2*M*N*K of each launch.
    peaks = load_peaks()
    unet.enable_cuda_graph(False)
    rec = []
    orig_conv = ops.conv

    def conv_timed(srcs, weight, out, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig_conv(srcs, weight, out, **kw)
        e1.record()
        wimg = kw.get("wimg")
        phased = bool(kw.get("upsample_phases"))
        tc = srcs[0].x.dtype == torch.bfloat16 and (wimg is not None or phased or
                                                     all(s.x.c % 64 == 0 and not s.upsample for s in srcs))
        # algorithmic bytes of the launch: every operand once -- A (each source read once, NOT once per filter tap),
        # W, bias / row vector, residual, output and its bf16 shadow
        nbytes = sum(s.x.rows * s.x.c * s.x.buf.element_size() for s in srcs)
        nbytes += out.rows * out.c * out.buf.element_size()
        if wimg is not None:
            # per-image B operand (collapsed cross-attention): [heads * tokens, C] per image, K = the source's channels
            _, tokens, heads, hstride = wimg
            k_eff, n_eff = srcs[0].x.c, out.c
            nbytes += srcs[0].x.n * tokens * heads * hstride * 2
        elif phased:
            # four 2x2 output phases: every output pixel multiplies 4 of the 16 packed tap blocks
            k_eff, n_eff = weight.shape[1] // 4, weight.shape[0]
            nbytes += weight.numel() * weight.element_size()
        else:
            k_eff, n_eff = weight.shape[1], weight.shape[0]
            nbytes += weight.numel() * weight.element_size()
        for key in ("residual", "out2"):
            t_ = kw.get(key)
            if t_ is not None:
                nbytes += t_.rows * t_.c * t_.buf.element_size()
        if kw.get("bias") is not None:
            nbytes += kw["bias"].numel() * 4
        if kw.get("rowvec") is not None:
            nbytes += srcs[0].x.n * weight.shape[0] * 4
        rec.append((e0, e1, 2.0 * out.rows * n_eff * k_eff, tc, nbytes, kw.get("gn_apply") is not None))
        return r

    xin = torch.cat([x_T_d] * 2)
    tin = torch.full((2 * B,), 981, device=dev, dtype=torch.long)
    cin = torch.cat([uc_d, cond_d])
    # the same forward the sampler issues: guidance pair (prefix in front of the first cross-attention computed once)
    # inside a sampling scope (context projected by the warm-up call, not again)
    with unet.sampling_scope(), unet.cfg_pair():
        unet(xin, tin, context=cin)          # eager warm-up
        l0 = ops.launch_count()
        ops.conv = conv_timed
        import ealdm_b200.unet as _unet_mod
        _unet_mod.ops.conv = conv_timed
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # a device-side delay first, so that the host enqueues the whole forward ahead of the GPU: the event pairs then
        # time back-to-back kernels instead of kernels + the host's launch latency
        torch.cuda._sleep(int(6e7))
        f0.record()
        unet(xin, tin, context=cin)
        f1.record()
        torch.cuda.synchronize()
        ops.conv = orig_conv
        _unet_mod.ops.conv = orig_conv
    launches_per_forward = ops.launch_count() - l0
    tc_ms = sum(r_[0].elapsed_time(r_[1]) for r_ in rec if r_[3])
    tc_flops = sum(r_[2] for r_ in rec if r_[3])
    tc_bytes = sum(r_[4] for r_ in rec if r_[3])
    # the launches whose epilogue also applies a GroupNorm (+ SiLU) do the work of a second kernel: reported apart
    gn_ms = sum(r_[0].elapsed_time(r_[1]) for r_ in rec if r_[3] and r_[5])
    gn_flops = sum(r_[2] for r_ in rec if r_[3] and r_[5])
    n_gn = sum(1 for r_ in rec if r_[3] and r_[5])
    n_tc = sum(1 for r_ in rec if r_[3])
    fwd_ms = f0.elapsed_time(f1)
    achieved = tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
    roofline = {"bound": "tensor", "kernel": "ealdm::tc::conv_tc_kernel (tcgen05 implicit-GEMM conv/linear)",
                "achieved": achieved, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                "frac": achieved / peaks["bf16_sustained"], "peak_source": peaks["source"] + " (sustained)",
                "traffic": None, "algorithmic_bytes": tc_bytes / max(n_tc, 1),
                "algorithmic_bytes_per_forward": tc_bytes, "algorithmic_flops_per_forward": tc_flops,
                "launches_per_forward": n_tc, "avg_launch_ms": tc_ms / max(n_tc, 1),
                "launches_with_groupnorm_epilogue": n_gn,
                "achieved_without_groupnorm_epilogue": ((tc_flops - gn_flops) / ((tc_ms - gn_ms) * 1e-3) / 1e12
                                                        if tc_ms > gn_ms else None),
                "achieved_with_groupnorm_epilogue": gn_flops / (gn_ms * 1e-3) / 1e12 if gn_ms > 0 else None,
                # share of one DDIM step of the timed (CUDA-graph) run; the instrumented eager forward itself is
                # host-bound (an event pair per launch), so its own duration is reported separately
                "share_of_forward": tc_ms / (ms_per_step / S) if ms_per_step > 0 else None,
                "eager_instrumented_forward_ms": fwd_ms,
                "note": "event-timed eager forward at UNet batch %d as the sampler issues it (guidance pair: the layers "
                        "in front of the first cross-attention run on half the batch; context projection made once per "
                        "sampling loop, not counted), enqueued behind a device-side delay so that the "
                        "launches run back to back; algorithmic FLOPs = 2*M*N*K per launch AS EXECUTED (upsampling "
                        "phases: K = 4C per output; collapsed cross-attention: its two small per-image GEMMs); "
                        "algorithmic_bytes = A + W + bias/rowvec + residual + out (+ bf16 shadow), every operand once, "
                        "per-launch average like `traffic`; share_of_forward = summed launch time / (ms_per_step / "
                        "ddim_steps)" % (2 * B)}
    # DRAM bytes per launch of the same kernel from the committed ncu capture of one forward at this batch
    tpath = os.path.join(ROOT, "profiles", "r02_conv_tc_traffic.json")
    if not os.path.exists(tpath):
        tpath = os.path.join(ROOT, "profiles", "r01_conv_tc_traffic.json")
    if os.path.exists(tpath) and B == 64:
        with open(tpath) as f:
            tj = json.load(f)
        roofline["traffic"] = tj["dram_bytes_per_launch"]
        roofline["traffic_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum averaged over the %d launches of one "
                                    "forward (%s); L2 traffic per launch %.0f MB"
                                    % (tj["launches"], os.path.relpath(tpath, ROOT), tj["l2_bytes_per_launch"] / 1e6))
    step_tflops = value / world * S * 2 * CFG.UNET_STDIFF_GFLOP_PER_SAMPLE / 1e3

    # --- extras of the default line (N = 1): the eager-PyTorch GPU bar and BASELINE.json configs[2] / configs[4] -----
    extras = {}
    if world == 1 and not args.no_extras and not args.ncu_window:
        unet.invalidate_packed()           # drops the packed weights and the CUDA graphs (their memory pool)
        torch.cuda.empty_cache()
        extras["torch_gpu_baseline"] = torch_gpu_baseline_leg(dev, B, S)
        sub = argparse.Namespace(**vars(args))
        sub.quiet, sub.steps, sub.warmup, sub.batch = True, 3, 3, 64
        for key, fn in (("config3_autoencoder", run_autoencoder), ("config5_train", run_train)):
            try:
                r = fn(sub)
                extras[key] = {k: r[k] for k in ("metric", "value", "unit", "ms_per_step", "e2e", "config", "gpu_launches",
                                                 "frac_of_sustained_peak") if k in r}
            except Exception as ex:
                extras[key] = {"error": f"{type(ex).__name__}: {ex}"}
            torch.cuda.empty_cache()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD if not strong else
                       (f"stdiff_cin-ldm-vq-f8 conditioned DDIM sampling with classifier-free guidance, FIXED global batch "
                        f"{Bg} sharded over {world} GPU(s) = {B} per GPU (BASELINE.json configs[3]), 50-step DDIM, CFG 2.0"),
                       "batch_per_gpu": B, "global_batch": Bg, "ddim_steps": S, "guidance_scale": ugs,
                       "unet_batch": 2 * B, "parallelism": f"batch-sharded x{world}, one final all-gather",
                       "cuda_graph": not args.no_graph,
                       "guidance_pair": "the cond / uncond halves share x and t: layers in front of the first "
                                        "cross-attention are computed once (bit-identical; EALDM_NO_CFG_SHARE=1 "
                                        "disables), context projected once per sampling loop",
                       "l2": "working set per step (0.79 GB bf16 weights + >4 GB activations per UNet forward) >> 126 MB L2",
                       "weights": "random-init, BASELINE.md section 4 distribution"},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e},
            "gpu_launches": int(args.steps * S * (launches_per_forward + 1)),
            "launches_per_unet_forward": int(launches_per_forward),
            "roofline": roofline,
            "unet_tflops_per_gpu": step_tflops,
            "unet_frac_of_sustained_peak": step_tflops / peaks["bf16_sustained"],
            "cpu_baseline": cpu,
            "with_decode": with_decode,
        }
        line.update(extras)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
