"""pea_diffusion_b200 -- B200-native (sm_100a) hot path of PEA-Diffusion's KD training:
PEA adapter (``adapter.MLP``), the UNet's cross-attention as a diffusers-protocol processor (``attn2``), and the
teacher/student KD loss (``kd_loss``), all on hand-written CUDA behind the C ABI of ``include/pea_b200.h``.

Importing the package loads ``libpea_b200.so``; there is no CPU / PyTorch fallback for the kernels.
"""
from . import _lib  # noqa: F401  (fails loudly when the CUDA library has not been built)
from .adapter import MLP, MLP_Legacy, MLP_SD, MLP_plus  # noqa: F401
from .attn2 import PEAAttn2Processor, attn2_core, install  # noqa: F401
from .kd_loss import kd_loss  # noqa: F401
from .staging import BatchStager  # noqa: F401

__version__ = "0.1.0"
