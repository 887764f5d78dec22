"""Event-timed forward of the EALDM conditioner (UnetCond on the VQ-f8 first stage's encoder) at the benchmark batch:
64 frames 256 x 256 -> context [64, 4, 512].  Prints the split encoder / rest and the launch count."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from ealdm_b200 import configs as CFG, ops  # noqa: E402
from ealdm_b200.autoencoder import VQModelInterface  # noqa: E402
from ealdm_b200.conditioner import UnetCond  # noqa: E402


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    torch.manual_seed(0)
    first = VQModelInterface(embed_dim=CFG.VQ_F8_EMBED_DIM, n_embed=CFG.VQ_F8_N_EMBED,
                             ddconfig=dict(CFG.VQ_F8_DDCONFIG)).cuda().eval()
    cond = UnetCond(cond_args={"type": "fourier", "f_manual": [1.839835728952772, 672], "include_lin": True,
                               "lin_lr": 0.01, "lr": 1, "dims": 6}).cuda().eval()
    cond.convs = first
    g = torch.Generator().manual_seed(1)
    mixed = (torch.rand(B, 3, 256, 256, generator=g).cuda() * 2 - 1, torch.rand(B, 1, 1, generator=g).cuda(),
             torch.randn(B, 1, 16, generator=g).cuda(), torch.rand(B, 1, generator=g).cuda())
    n0 = ops.launch_count()
    cond(mixed)
    launches = ops.launch_count() - n0
    ms_all = timed(lambda: cond(mixed))
    ms_enc = timed(lambda: first._eng(mixed[0]).encoder_features(mixed[0]))
    print(f"conditioner forward, {B} frames: {ms_all:.2f} ms ({launches} launches): first-stage encoder {ms_enc:.2f} ms "
          f"(bf16, tcgen05), styles + AdaIN + conv_cat + out_layer {ms_all - ms_enc:.2f} ms (fp32)")


if __name__ == "__main__":
    main()
