"""CPU tests of the host side: plugin boundary, schedule bit-exactness of the product code, state-dict
compatibility, C-ABI exports, repository layout rules, and the multi-process sharding logic (gloo)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
PKG = os.path.join(ROOT, "environment-aware_latent_diffusion_model_b200")

from ealdm_b200 import _lib as L  # noqa: E402
from ealdm_b200 import configs as CFG  # noqa: E402
from ealdm_b200.ddim import DDIMSampler  # noqa: E402
from ealdm_b200.ddpm import LatentDiffusion  # noqa: E402
from ealdm_b200.util import instantiate_from_config  # noqa: E402
from oracle import autoencoder as OA  # noqa: E402
from oracle import unet as OU  # noqa: E402


def gold(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def test_abi_library_exports_every_declared_symbol():
    """No compute calls (no GPU here): the .so loads and exports what include/ealdm_b200.h declares."""
    L.build()
    lib = L.load()
    with open(os.path.join(ROOT, "include", "ealdm_b200.h")) as f:
        declared = sorted(set(re.findall(r"\b(ealdm_[a-z0-9_]+)\s*\(", f.read())))
    assert len(declared) >= 18
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert sorted(L.EXPORTS) == declared
    assert lib.ealdm_abi_version() == 2
    # argument validation happens before any CUDA call
    a = L.LayerNormArgs()
    assert lib.ealdm_layer_norm(ctypes.byref(a), None) == -1
    assert b"null" in lib.ealdm_last_error()
    # per-(image, pixel chunk, 4-channel vector) float2 partials; 2368 CTAs / 128 images -> 19 chunks
    assert lib.ealdm_group_norm_workspace_bytes(128, 1024, 256) == 128 * 19 * 64 * 8


def test_struct_layout_matches_the_header():
    src = '#include <stdio.h>\n#include "ealdm_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n",' \
          "sizeof(ealdm_conv_src),sizeof(ealdm_conv_args),sizeof(ealdm_group_norm_args),sizeof(ealdm_layer_norm_args)," \
          "sizeof(ealdm_attention_args),sizeof(ealdm_ddim_step_args));return 0;}"
    exe = "/tmp/ealdm_sizeof"
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=src.encode(), check=True)
    sizes = [int(v) for v in subprocess.run([exe], capture_output=True, check=True).stdout.split()]
    assert sizes == [ctypes.sizeof(c) for c in (L.ConvSrc, L.ConvArgs, L.GroupNormArgs, L.LayerNormArgs,
                                                L.AttentionArgs, L.DdimStepArgs)]


def test_product_never_imports_the_oracle_and_has_no_cpu_fallback():
    for dirpath, _, files in os.walk(PKG):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), fn
    from ealdm_b200.unet import UNetModel
    m = UNetModel(**CFG.UNET_UNCOND)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 4, 32, 32), torch.zeros(1, dtype=torch.long))


def test_instantiate_from_config_yaml_targets():
    for rel, key in (("configs/latent-diffusion/stdiff_cin-ldm-kl-f8_b200.yaml", 626),
                     ("configs/latent-diffusion/uncond_cin-ldm-vq-f8_b200.yaml", 306)):
        cfg = yaml.safe_load(open(os.path.join(ROOT, rel)))["model"]
        model = instantiate_from_config(cfg)
        assert type(model).__name__ == "LatentDiffusion"
        unet = model.model.diffusion_model
        assert len(unet.state_dict()) == key
        assert unet.in_channels == 4 and unet.image_size == 32
    ae = instantiate_from_config(yaml.safe_load(open(os.path.join(ROOT, "configs/autoencoder/autoencoder_kl_32x32x4_b200.yaml")))["model"])
    assert hasattr(ae, "encode") and hasattr(ae, "decode")


def test_shipped_stdiff_config_wires_conditioner_to_first_stage():
    """configs/latent-diffusion/stdiff_cin-ldm-vq-f8_b200.yaml = the reference's shipped config with all four targets
    switched: VQ first stage, UnetCond conditioner whose `convs` IS the first stage (ddpm.py:535-536), same values."""
    from oracle import conditioner as OC
    cfg = yaml.safe_load(open(os.path.join(ROOT, "configs/latent-diffusion/stdiff_cin-ldm-vq-f8_b200.yaml")))["model"]
    model = instantiate_from_config(cfg)
    assert type(model.first_stage_model).__name__ == "VQModelInterface"
    cond = model.cond_stage_model
    assert type(cond).__name__ == "UnetCond" and cond.convs is model.first_stage_model
    own = [(k, tuple(v.shape)) for k, v in cond.state_dict().items() if not k.startswith("convs.")]
    assert own == OC.param_shapes()
    assert dict(cond.cond_args)["f_manual"] == OC.COND_ARGS["f_manual"] and cond.cond_args.lin_lr == 0.01
    with pytest.raises(RuntimeError):          # no CPU fallback
        cond((torch.zeros(2, 3, 8, 8), torch.zeros(2, 1, 1), torch.zeros(2, 1, 16), torch.zeros(2, 1)))


@pytest.mark.parametrize("cfg", [CFG.UNET_STDIFF, CFG.UNET_UNCOND])
def test_unet_state_dict_matches_reference_inventory(cfg):
    from ealdm_b200.unet import UNetModel
    m = UNetModel(**cfg)
    mine = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    assert mine == OU.unet_param_shapes(cfg)          # same names, shapes AND registration order
    sd = OU.synthetic_state_dict(OU.unet_param_shapes(cfg), seed=5)
    m.load_state_dict(sd, strict=True)
    # the reference zero-initialises these (openaimodel.py:229-231,312,685; attention.py:244-248)
    fresh = UNetModel(**cfg)
    zeros = [k for k, v in fresh.state_dict().items() if v.dim() >= 2 and float(v.abs().max()) == 0.0]
    assert len(zeros) == 34
    for combo in (dict(use_scale_shift_norm=True), dict(resblock_updown=True), dict(num_classes=10), dict(dims=3)):
        with pytest.raises(NotImplementedError):
            UNetModel(**{**cfg, **combo})


def test_autoencoder_state_dict_matches_reference_inventory():
    from ealdm_b200.autoencoder import AutoencoderKL
    ae = AutoencoderKL(ddconfig=dict(CFG.AE_KL_F8_DDCONFIG), embed_dim=4)
    mine = [(k, tuple(v.shape)) for k, v in ae.state_dict().items()]
    assert mine == OA.autoencoder_kl_param_shapes(CFG.AE_KL_F8_DDCONFIG, 4)


def test_vq_first_stage_state_dict_matches_reference_inventory():
    from ealdm_b200.autoencoder import VQModelInterface
    m = VQModelInterface(embed_dim=CFG.VQ_F8_EMBED_DIM, n_embed=CFG.VQ_F8_N_EMBED, ddconfig=dict(CFG.VQ_F8_DDCONFIG))
    mine = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    assert mine == OA.vq_param_shapes(CFG.VQ_F8_DDCONFIG, CFG.VQ_F8_EMBED_DIM, CFG.VQ_F8_N_EMBED)
    with pytest.raises(RuntimeError):
        m.decode(torch.zeros(1, 4, 8, 8))        # CPU tensor: no fallback


def test_schedules_bit_exact_vs_reference_golden():
    G = gold("schedule.pt")
    ld = LatentDiffusion(unet_config={"target": "ealdm_b200.unet.UNetModel", "params": dict(CFG.UNET_UNCOND)},
                         **CFG.DIFFUSION)
    for k, v in G["register"].items():
        assert torch.equal(getattr(ld, k), v), k
    for S in (10, 50):
        for eta in (0.0, 1.0):
            ref = G[f"S{S}_eta{eta}"]
            s = DDIMSampler(ld)
            s.make_schedule(S, ddim_eta=eta, verbose=False)
            assert torch.equal(torch.as_tensor(s.ddim_timesteps), ref["ddim_timesteps"])
            for k in ("ddim_alphas", "ddim_alphas_prev", "ddim_sigmas", "ddim_sqrt_one_minus_alphas"):
                assert torch.equal(torch.as_tensor(np.asarray(getattr(s, k)), dtype=torch.float64), ref[k]), (S, eta, k)
            # per-step fp32 scalars == what the reference's torch.full tensors hold
            for index in (0, S // 2, S - 1):
                sc = s.step_scalars(index)
                a_t = torch.full((), float(ref["ddim_alphas"][index]), dtype=torch.float32)
                a_p = torch.full((), float(ref["ddim_alphas_prev"][index]), dtype=torch.float32)
                sg = torch.full((), float(ref["ddim_sigmas"][index]), dtype=torch.float32)
                assert sc["sqrt_at"] == float(a_t.sqrt()) and sc["sqrt_a_prev"] == float(a_p.sqrt())
                assert sc["dir_coef"] == float((1. - a_p - sg ** 2).sqrt()) and sc["sigma_t"] == float(sg)
    s = DDIMSampler(ld)
    s.make_schedule(50, ddim_eta=1.0, verbose=False)
    assert list(s.ddim_timesteps[:2]) == [1, 21] and s.ddim_timesteps[-1] == 981


def test_geglu_interleave_roundtrip():
    import torch.nn.functional as F
    from ealdm_b200.packing import geglu_interleave, pack_conv_weight
    C = 64
    w, b, x = torch.randn(8 * C, C), torch.randn(8 * C), torch.randn(7, C)
    wp, bp = geglu_interleave(w, b)
    y = F.linear(x, wp, bp).reshape(7, -1, 2, 16)
    out = (y[:, :, 0] * F.gelu(y[:, :, 1])).reshape(7, 4 * C)
    val, gate = F.linear(x, w, b).chunk(2, -1)
    assert torch.allclose(out, val * F.gelu(gate), atol=1e-5)
    cw = torch.randn(8, 4, 3, 3)
    pw = pack_conv_weight(cw, torch.float32)
    assert pw.shape == (8, 36) and torch.equal(pw[:, 4 * (3 * 1 + 2):4 * (3 * 1 + 2) + 4], cw[:, :, 1, 2])


def test_shard_bounds_cover_batch_exactly():
    from ealdm_b200.parallel import shard_bounds
    for total in (0, 1, 7, 64, 512, 513):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    from ealdm_b200.parallel import sample_sharded
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(3)
    x_T = torch.randn(7, 4, 8, 8, generator=g)       # ragged: 7 samples over 2 ranks
    cond = torch.randn(7, 4, 16, generator=g)
    fake = lambda x, c, u: x * 2 + c.mean(dim=(1, 2))[:, None, None, None]  # noqa: E731  per-sample function
    out = sample_sharded(fake, x_T, cond, None)
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_sampling_gloo_world2_matches_single_process():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(3)
    x_T = torch.randn(7, 4, 8, 8, generator=g)
    cond = torch.randn(7, 4, 16, generator=g)
    ref = x_T * 2 + cond.mean(dim=(1, 2))[:, None, None, None]
    assert torch.equal(outs[0], ref) and torch.equal(outs[1], ref)


TINY_UNET = dict(image_size=8, in_channels=4, model_channels=32, out_channels=4, num_res_blocks=1,
                 attention_resolutions=[1, 2], channel_mult=(1, 2), num_head_channels=32,
                 use_spatial_transformer=True, transformer_depth=1, context_dim=64)


def _grad_worker(rank, world, port, q):
    import torch.distributed as dist
    from ealdm_b200.parallel import GradBuckets
    from ealdm_b200.unet import UNetModel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    unet = UNetModel(**TINY_UNET)
    gb = GradBuckets(unet, bucket_mb=0.05)          # many small buckets
    assert len(gb.bounds) > 4
    # every parameter got a view of the flat buffer, exactly once
    assert sum((p.numel() + 7) // 8 * 8 for p in unet.parameters()) == gb.flat.numel()
    for _ in range(2):                              # two "steps": zero_() must re-arm the buckets
        gb.zero_()
        for i, p in enumerate(unet.parameters()):
            assert p.grad.data_ptr() >= gb.flat.data_ptr()
            p.grad.add_(float(i + 1) * (rank + 1))
        # the order UNetTrainEngine.backward reports blocks in
        gb.block_done("head", 0)
        for j in reversed(range(len(unet.output_blocks))):
            gb.block_done("out", j)
        gb.block_done("mid", 0)
        launched_before_inputs = gb._next
        for i in reversed(range(len(unet.input_blocks))):
            gb.block_done("in", i)
        gb.finish()
    vals = [float(p.grad.flatten()[0]) for p in unet.parameters()]
    q.put((rank, vals, launched_before_inputs, len(gb.bounds)))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_buckets_gloo_world2_average_and_overlap_order():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, vals, early, nb in outs:
        # mean over ranks of (i+1)*(rank+1) = 1.5*(i+1)
        assert vals == [1.5 * (i + 1) for i in range(len(vals))]
        assert 0 < early < nb      # buckets were launched while "backward" was still running


def test_upsample_phase_weights_reproduce_nearest2x_conv3x3():
    """packing.pack_upsample_phases: four 2x2 phases over the low-resolution input == conv3x3(pad 1) over the
    nearest-2x upsampled input (openaimodel.py:109-119), borders included."""
    import torch.nn.functional as F
    from ealdm_b200.packing import pack_upsample_phases
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 5, 6, 7, generator=g, dtype=torch.float64)
    w = torch.randn(4, 5, 3, 3, generator=g, dtype=torch.float64)
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, padding=1)
    wp = pack_upsample_phases(w.float(), torch.float32).double().reshape(4, 2, 2, 2, 2, 5)
    xp = F.pad(x, (1, 1, 1, 1))
    out = torch.empty_like(ref)
    for py in (0, 1):
        for px in (0, 1):
            k = wp[:, py, px].permute(0, 3, 1, 2)
            out[:, :, py::2, px::2] = F.conv2d(xp[:, :, py:py + 7, px:px + 8], k)
    assert float((out - ref).abs().max()) < 1e-5


def test_collapsed_cross_attention_weights_reproduce_the_reference_layer():
    """packing.collapse_cross_attention: with G_h = scale log2(e) Wk_h^T Wq_h and H_h = Wv_h^T Wo_h^T the layer
    to_q -> softmax(q k^T scale) v -> to_out (CrossAttention.forward, ldm/modules/attention.py:170-193) is
    softmax_2(x (ctx G)^T) (ctx H) + b.  Checked in fp64 against the plain q / k / v evaluation, with the row order
    (head, key) / (head, channel) the per-image TMA gather of ealdm_conv(wi_*) assumes."""
    from ealdm_b200.packing import collapse_cross_attention
    g = torch.Generator().manual_seed(11)
    n, tok, T, heads, d, E = 2, 24, 4, 8, 32, 96
    C = heads * d
    dd = torch.float64
    x = torch.randn(n, tok, C, generator=g, dtype=dd)
    ctx = torch.randn(n, T, E, generator=g, dtype=dd)
    wq, wo = (torch.randn(C, C, generator=g, dtype=dd) / C ** 0.5 for _ in range(2))
    wk, wv = (torch.randn(C, E, generator=g, dtype=dd) / E ** 0.5 for _ in range(2))
    bo = torch.randn(C, generator=g, dtype=dd)
    scale = d ** -0.5
    q = (x @ wq.t()).reshape(n, tok, heads, d).transpose(1, 2)
    k = (ctx @ wk.t()).reshape(n, T, heads, d).transpose(1, 2)
    v = (ctx @ wv.t()).reshape(n, T, heads, d).transpose(1, 2)
    p = torch.softmax(q @ k.transpose(-1, -2) * scale, dim=-1)
    want = (p @ v).transpose(1, 2).reshape(n, tok, C) @ wo.t() + bo
    G, H = collapse_cross_attention(wq, wk, wv, wo, heads, scale, torch.float32)
    G, H = G.to(dd), H.to(dd)                                         # rows (head, channel)
    U = (ctx @ G.t()).reshape(n, T, heads, C).permute(0, 2, 1, 3)     # [n, head, key, c]: B rows (head, key) of image n
    Z = (ctx @ H.t()).reshape(n, T, heads, C).permute(0, 2, 1, 3)
    logits2 = torch.einsum("ntc,nhjc->nthj", x, U)                    # base-2 logits (log2 e folded into G)
    p2 = torch.softmax(logits2 * 0.6931471805599453, dim=-1)
    got = torch.einsum("nthj,nhjc->ntc", p2, Z) + bo
    assert float((got - want).norm() / want.norm()) < 1e-6            # G / H were rounded to fp32


def test_guidance_pair_and_sampling_scope_host_logic():
    """util.GuidancePair / UNetModel.sampling_scope (the hints the samplers give the UNet): the concatenated conditioning
    is built once while `uc` and `c` stay the same tensors at the same version, rebuilt when either changes, the hosted
    UNet sees the guidance-pair flag only inside the call, and a model without the hooks is driven exactly like the
    reference drives it (ddim.py:173-179)."""
    import contextlib

    from ealdm_b200.util import GuidancePair, sampling_scope

    class FakeUNet:
        def __init__(self):
            self.pair, self.scope, self.seen = False, False, []

        @contextlib.contextmanager
        def cfg_pair(self):
            self.pair = True
            try:
                yield self
            finally:
                self.pair = False

        @contextlib.contextmanager
        def sampling_scope(self):
            self.scope = True
            try:
                yield self
            finally:
                self.scope = False

    class Holder:
        pass

    class FakeModel:
        def __init__(self, unet):
            self.model = Holder()
            self.model.diffusion_model = unet
            self.calls = []

        def apply_model(self, x, t, c):
            u = self.model.diffusion_model
            self.calls.append((x, t, c, getattr(u, "pair", None), getattr(u, "scope", None)))
            return torch.cat([x[: x.shape[0] // 2] * 0 + 1, x[x.shape[0] // 2:] * 0 + 2])

    unet = FakeUNet()
    model = FakeModel(unet)
    pair = GuidancePair(model)
    x, t = torch.randn(3, 4, 2, 2), torch.tensor([5, 5, 5])
    uc, c = torch.randn(3, 4, 8), torch.randn(3, 4, 8)
    with sampling_scope(model):
        e_u, e_c = pair(x, t, uc, c)
        pair(x + 1, t, uc, c)
    assert torch.equal(e_u, torch.ones_like(x)) and torch.equal(e_c, 2 * torch.ones_like(x))
    (x0, t0, c0, p0, s0), (x1, t1, c1, p1, s1) = model.calls
    assert torch.equal(x0, torch.cat([x, x])) and torch.equal(t0, torch.cat([t, t])) and torch.equal(c0, torch.cat([uc, c]))
    assert p0 and s0 and p1 and s1 and not unet.pair and not unet.scope      # flags only inside the calls
    assert c1 is c0                                                          # same uc / c objects: the same c_in tensor
    c.add_(1.0)                                                              # in-place change: version bump
    pair(x, t, uc, c)
    assert model.calls[2][2] is not c0 and torch.equal(model.calls[2][2], torch.cat([uc, c]))
    uc2 = uc.clone()
    pair(x, t, uc2, c)
    assert model.calls[3][2] is not model.calls[2][2]
    # a model that does not host the B200 UNet: no hooks, same calls
    plain = FakeModel(object())
    with sampling_scope(plain):
        GuidancePair(plain)(x, t, uc, c)
    assert plain.calls[0][3] is None and torch.equal(plain.calls[0][2], torch.cat([uc, c]))


def test_activation_views_follow_the_statistics_buffer():
    """ops.Act.images / cols: views of an NHWC activation share its storage, and the GroupNorm partial-statistics buffer
    (one row per 32-pixel chunk) is sliced with the images (the guidance-pair prefix runs on the first half of a batch)."""
    from ealdm_b200.ops import Act
    a = Act(torch.arange(4 * 8 * 8 * 64, dtype=torch.float32).reshape(4 * 64, 64), 4, 8, 8)
    a.gp = torch.arange(4 * 2 * 8 * 2, dtype=torch.float32).reshape(4 * 2, 8, 2)     # 2 chunks per image, 8 octets
    v = a.images(1, 2)
    assert (v.n, v.h, v.w, v.c, v.c0, v.rows) == (2, 8, 8, 64, 0, 128)
    assert v.buf.data_ptr() == a.buf[64:].data_ptr() and v.gp.data_ptr() == a.gp[2:].data_ptr() and v.gp.shape[0] == 4
    w = a.cols(32, 16).images(2, 1)
    assert (w.c0, w.c, w.n) == (32, 16, 1) and torch.equal(w.view2d(), a.buf[128:192, 32:48])
    w.view2d().zero_()
    assert float(a.buf[128:192, 32:48].abs().sum()) == 0.0 and float(a.buf[128:192, :32].abs().sum()) > 0
    with pytest.raises(AssertionError):
        a.images(3, 2)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm): one JSON line with the same metric /
    unit / config keys as the GPU arm plus impl, cpu_baseline {kind, cores, sample, value} and an e2e object with zero
    copy bytes; it must run without a GPU and without /root/reference."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"].startswith("latent samples/sec") and line["unit"] == "samples/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 1
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == line["value"] > 0
    e = line["e2e"]
    assert e["value"] == line["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and line["config"].get("same_config") is False
