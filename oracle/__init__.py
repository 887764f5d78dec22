"""CPU oracle for the EALDM denoising hot path -- TEST INFRASTRUCTURE ONLY.

A plain-PyTorch (fp32, CPU) functional restatement of the reference's arithmetic for the path named
in BASELINE.json's north_star: UNetModel forward, DDIM sampling / q_sample / p_losses and the
AutoencoderKL encoder/decoder.  Every function cites the reference file:line it follows
(paths relative to the reference repository root, NasrinKalanat/Environment-Aware_Latent_Diffusion_Model).

Rules (checked by tests/test_layout.py):
  * only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg
    may import this package -- as the checker, never as the thing measured or shipped;
  * the product package must never import it and has no CPU fallback.

Pinning: the reference ships NO tests or golden vectors for this path, so the
oracle is pinned against outputs of the reference itself: `oracle/gen_golden.py` imports the
reference's own modules from /root/reference (with import shims for the absent pytorch_lightning /
omegaconf / taming packages), runs them on seeded inputs and commits the results under
`tests/golden/`; `tests/test_oracle_golden.py` checks this restatement against those files.
"""
