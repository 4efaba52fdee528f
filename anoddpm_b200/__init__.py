"""anoddpm_b200 -- B200-native (sm_100a) implementation of AnoDDPM's hot path: the Gaussian-diffusion
sampling loop, the UNet denoiser it calls each step and the simplex-noise generator, behind the
reference's own Python API (modules GaussianDiffusion / UNet / simplex; put ``dropin/`` first on
sys.path to import them under the reference's names).

Everything numeric runs in hand-written CUDA kernels behind a C ABI (include/anoddpm_b200.h,
anoddpm_b200/libanoddpm_b200.so).  There is no CPU or PyTorch fallback on the inference path.
"""
from . import _lib  # noqa: F401
from .GaussianDiffusion import GaussianDiffusionModel, get_beta_schedule  # noqa: F401
from .UNet import UNetModel, update_ema_params  # noqa: F401
from .simplex import Simplex_CLASS  # noqa: F401

__version__ = "0.1.0"
