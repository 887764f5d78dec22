"""Importable alias of the `environment-aware_latent_diffusion_model_b200` package (whose directory
name is not a Python identifier).  `import ealdm_b200.unet` == the module of the same name there."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_REAL = "environment-aware_latent_diffusion_model_b200"
_pkg = importlib.import_module(_REAL)
__path__ = _pkg.__path__          # submodule imports resolve inside the real package directory
__version__ = _pkg.__version__


def real_name() -> str:
    return _REAL
