"""TEST INFRASTRUCTURE ONLY -- golden GRADIENTS of the EALDM conditioner: the reference
trains `UnetCond` jointly with the UNet (`cond_stage_trainable: true`, ldm/models/diffusion/ddpm.py:1409-1415), so the
B200 module needs a backward.  torch.autograd through the reference's own `STDiff.models.UnetCond`
(STDiff/models.py:411-539), with the frozen first-stage encoder replaced by a stub that returns the golden encoder
output z of tests/golden/conditioner.pt (the encoder is frozen: no gradient flows into or through it), for the scalar
    L = sum(context * R),  R ~ N(0, 1) seeded,
in eval mode and with BatchNorm on batch statistics (Dropout off in both: its mask is RNG plumbing), plus the
negative-conditioning branch (mixed[-1] is None: out_layer only) and a 4-step WeatherLSTM recurrence on its own.

-> tests/golden/conditioner_grads.pt.  Run in the build container:  python oracle/gen_golden_cond_grads.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

SEED_R, SEED_RNEG, SEED_RSEQ = 96, 97, 98
FULL_MAX = 70_000          # larger gradient tensors are stored as {norm, 4 random projections, a strided sample}
SAMPLE_STRIDE = 997


def direction(name: str, k: int, shape):
    """Seeded random direction k for the tensor called `name` (shared with tests/test_conditioner_gpu.py)."""
    seed = (sum(ord(ch) * (i + 1) for i, ch in enumerate(name)) * 31 + k) % (2 ** 31)
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))


def compress(name: str, g: torch.Tensor):
    if g.numel() <= FULL_MAX:
        return {"full": g.clone()}
    gd = g.double()
    return {"norm": float(gd.norm()), "proj": [float((gd * direction(name, k, g.shape).double()).sum()) for k in range(4)],
            "sample": g.reshape(-1)[::SAMPLE_STRIDE].clone()}


class _Args(dict):
    __getattr__ = dict.get


def main():
    import gen_golden as GG
    GG.install_shims()
    from oracle import conditioner as OC
    torch.cuda.current_device = lambda: "cpu"
    import torchvision
    _resnet50 = torchvision.models.resnet50
    torchvision.models.resnet50 = lambda pretrained=False, **k: _resnet50(weights=None)
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    from STDiff.models import UnetCond

    G = torch.load(os.path.join(ROOT, "tests", "golden", "conditioner.pt"), weights_only=False)
    T = G["T"]
    m = UnetCond(device="cpu", cond_args=_Args(OC.COND_ARGS))

    class _Given(torch.nn.Module):          # stands for the frozen first stage: encoder(img) == golden z
        def encoder(self, img):
            return G["z"].clone()

    m.convs = _Given()
    missing, unexpected = m.load_state_dict(OC.synthetic_state_dict(), strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    own = [(n, p) for n, p in m.named_parameters()]
    assert [n for n, _ in own] == [n for n, s in OC.param_shapes() if "running" not in n and "num_batches" not in n]

    dummy = torch.zeros(T, 3, 8, 8)
    R = torch.randn(T, 4, 512, generator=torch.Generator().manual_seed(SEED_R))
    out = {"seeds": {"R": SEED_R, "R_neg": SEED_RNEG, "R_seq": SEED_RSEQ}}

    def grads(bn_train, mixed, Rw):
        m.eval()
        if bn_train:
            m.conv_cat[1].train()
        for _, p in own:
            p.grad = None
        bn = m.conv_cat[1]
        keep = (bn.running_mean.clone(), bn.running_var.clone(), bn.num_batches_tracked.clone())
        ctx = m(mixed)
        (ctx * Rw).sum().backward()
        bn.running_mean.copy_(keep[0]); bn.running_var.copy_(keep[1]); bn.num_batches_tracked.copy_(keep[2])
        return {n: compress(n, torch.zeros_like(p) if p.grad is None else p.grad) for n, p in own}, ctx.detach().clone()

    mixed = (dummy, G["flow"], G["weather"], G["time"])
    out["eval"], out["context_eval"] = grads(False, mixed, R)
    out["bn_train"], out["context_bn_train"] = grads(True, mixed, R)
    Rn = torch.randn(T, 4, 512, generator=torch.Generator().manual_seed(SEED_RNEG))
    out["negative"], out["context_negative"] = grads(False, [dummy, G["flow"], G["weather"], G["time"], None, None, None, None], Rn)
    # WeatherLSTM alone over a 4-step sequence: back-propagation through time
    seq = G["lstm_seq_in"]
    Rs = torch.randn(seq.shape[0] * seq.shape[1], 128, generator=torch.Generator().manual_seed(SEED_RSEQ))
    for _, p in own:
        p.grad = None
    m.eval()
    (m.w_mlp(seq, "train") * Rs).sum().backward()
    out["lstm_seq"] = {n: compress(n, p.grad) for n, p in own if n.startswith("w_mlp.")}
    path = os.path.join(ROOT, "tests", "golden", "conditioner_grads.pt")
    torch.save(out, path)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")
    for k in ("eval", "bn_train", "negative"):
        nrm = {n: (float(g["full"].norm()) if "full" in g else g["norm"]) for n, g in out[k].items()}
        print(f"{k}: {sum(1 for v in nrm.values() if v > 0)} of {len(nrm)} tensors receive a gradient; "
              f"|grad| total = {sum(v ** 2 for v in nrm.values()) ** 0.5:.4e}")
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())  # noqa: E731
    print("context reproduces conditioner.pt: eval %.1e, bn-train %.1e" % (
        rel(out["context_eval"], G["context_eval"]), rel(out["context_bn_train"], G["context_bn_train"])))


if __name__ == "__main__":
    main()
