"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the vector quantiser of the reference's VQ first stage.

PARITY UNPINNED: the algorithm lives in a third-party dependency that is NOT under /root/reference:
taming-transformers `taming/modules/vqvae/quantize.py: VectorQuantizer2`, installed by the reference as
`-e git+https://github.com/CompVis/taming-transformers.git@master` (environment.yaml:25, no pinned revision).  Its
published forward pass (legacy=True, remap=None, the arguments of the reference's call site
ldm/models/autoencoder.py:39-41) is restated below; parity is anchored on the reference's own call sites
(`VQModelInterface.decode`, autoencoder.py:274-282: quantize -> post_quant_conv -> decoder) by running the reference's
VQModelInterface with this restatement plugged in as `taming.modules.vqvae.quantize.VectorQuantizer2`
(oracle/gen_golden_vq.py).
"""
from typing import Tuple

import torch
import torch.nn as nn


def vq_nearest(z: torch.Tensor, codebook: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """VectorQuantizer2.forward, inference part:
        z = rearrange(z, 'b c h w -> b h w c'); z_flattened = z.view(-1, e_dim)
        d = sum(z_flattened^2, 1, keepdim) + sum(E^2, 1) - 2 * einsum('bd,dn->bn', z_flattened, E^T)
        min_encoding_indices = argmin(d, 1); z_q = E[indices].view(z.shape)
        z_q = z + (z_q - z).detach(); z_q = rearrange(z_q, 'b h w c -> b c h w')
    Returns (z_q NCHW, indices [b*h*w])."""
    zp = z.permute(0, 2, 3, 1).contiguous()
    zf = zp.view(-1, codebook.shape[1])
    d = torch.sum(zf ** 2, dim=1, keepdim=True) + torch.sum(codebook ** 2, dim=1) - \
        2 * torch.einsum("bd,dn->bn", zf, codebook.t())
    idx = torch.argmin(d, dim=1)
    zq = codebook[idx].view(zp.shape)
    zq = zp + (zq - zp).detach()
    return zq.permute(0, 3, 1, 2).contiguous(), idx


class VectorQuantizer2(nn.Module):
    """Stand-in with the constructor the reference calls (autoencoder.py:39-41) and taming's return convention."""

    def __init__(self, n_e, e_dim, beta, remap=None, unknown_index="random", sane_index_shape=False, legacy=True):
        super().__init__()
        assert remap is None
        self.n_e, self.e_dim, self.beta, self.legacy, self.sane_index_shape = n_e, e_dim, beta, legacy, sane_index_shape
        self.embedding = nn.Embedding(n_e, e_dim)
        self.embedding.weight.data.uniform_(-1.0 / n_e, 1.0 / n_e)

    def forward(self, z, temp=None, rescale_logits=False, return_logits=False):
        zq, idx = vq_nearest(z, self.embedding.weight)
        zp = z.permute(0, 2, 3, 1)
        zqp = zq.permute(0, 2, 3, 1)
        if not self.legacy:
            loss = self.beta * torch.mean((zqp.detach() - zp) ** 2) + torch.mean((zqp - zp.detach()) ** 2)
        else:
            loss = torch.mean((zqp.detach() - zp) ** 2) + self.beta * torch.mean((zqp - zp.detach()) ** 2)
        if self.sane_index_shape:
            idx = idx.reshape(z.shape[0], z.shape[2], z.shape[3])
        return zq, loss, (None, None, idx)
