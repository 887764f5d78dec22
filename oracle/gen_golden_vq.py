"""TEST INFRASTRUCTURE ONLY -- golden vectors of the VQ first stage, made by the
reference's own `VQModelInterface` (ldm/models/autoencoder.py:263-282: its Encoder / Decoder / quant convs, vq-f8
ddconfig of configs/latent-diffusion/stdiff_cin-ldm-vq-f8.yaml:37-60) with the restated taming quantiser
(oracle/vq.py, parity unpinned) plugged in for the un-vendored `taming.modules.vqvae.quantize.VectorQuantizer2`.
-> tests/golden/vq_f8.pt: a latent h [1,4,32,32], the quantiser's indices, decode(h) with and without quantisation.
Run in the build container (needs /root/reference):  python oracle/gen_golden_vq.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    import gen_golden as GG
    GG.install_shims()
    from oracle import autoencoder as OA
    from oracle import unet as OU
    from oracle import vq as OV
    sys.modules["taming.modules.vqvae.quantize"].VectorQuantizer2 = OV.VectorQuantizer2
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    from ldm.models.autoencoder import VQModelInterface

    cfg = GG.load_cfg("configs/latent-diffusion/stdiff_cin-ldm-vq-f8.yaml")["model"]["params"]["first_stage_config"]["params"]
    dd, n_embed, embed_dim = cfg["ddconfig"], cfg["n_embed"], cfg["embed_dim"]
    m = VQModelInterface(embed_dim=embed_dim, n_embed=n_embed, ddconfig=dd, lossconfig={"target": "torch.nn.Identity"})
    shapes = OA.vq_param_shapes(dd, embed_dim, n_embed)
    ref_names = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    assert ref_names == shapes, "vq-f8: oracle parameter inventory differs from the reference"
    sd = OU.synthetic_state_dict(shapes, seed=4)
    # a codebook that actually covers the latent range (uniform(-1/n, 1/n) would map everything near zero)
    sd["quantize.embedding.weight"] = torch.randn(n_embed, embed_dim, generator=GG.g(81)) * 1.2
    m.load_state_dict(sd, strict=True)
    m.eval()
    h = torch.randn(1, 4, 32, 32, generator=GG.g(82))
    img = torch.rand(1, 3, 64, 64, generator=GG.g(83)) * 2 - 1
    with torch.no_grad():
        _, _, (_, _, idx) = m.quantize(h)
        dec_q = m.decode(h)
        dec_nq = m.decode(h, force_not_quantize=True)
        enc = m.encode(img)
    path = os.path.join(ROOT, "tests", "golden", "vq_f8.pt")
    torch.save({"h": h, "indices": idx, "dec": dec_q, "dec_noquant": dec_nq, "img": img, "enc": enc}, path)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
