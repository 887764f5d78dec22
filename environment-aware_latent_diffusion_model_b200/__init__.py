"""ealdm-b200: Blackwell-native (sm_100a) implementation of EALDM's latent-diffusion denoising hot path.

Host code mirrors the reference's module/plugin interface (`UNetModel`, `AutoencoderKL`,
`DDIMSampler`, `LatentDiffusion`, `instantiate_from_config`); all device arithmetic goes through the
C ABI of `libealdm_b200.so` (see include/ealdm_b200.h).  The directory name contains a hyphen, so
import it with `importlib.import_module("environment-aware_latent_diffusion_model_b200")`, through a
config `target:` string, or through the `ealdm_b200` alias package at the repository root.
"""
__version__ = "0.1.0"
