"""pea_diffusion_b200 -- B200-native (sm_100a) hot path of PEA-Diffusion's KD training:
PEA adapter (``adapter.MLP``), the UNet's cross-attention as a diffusers-protocol processor (``attn2``), and the
teacher/student KD loss (``kd_loss``), all on hand-written CUDA behind the C ABI of ``include/pea_b200.h``.

The product modules bind ``libpea_b200.so`` the moment any of them is touched (``from pea_diffusion_b200 import MLP``,
``import pea_diffusion_b200.ops`` ...) and raise if it has not been built; there is no CPU / PyTorch fallback for the
kernels.  Only ``pea_diffusion_b200.shell`` -- the plain-PyTorch UNet stand-in that HOSTS the path in tests and
benchmarks -- can be imported without the library, so that the CPU reference arm (``bench.py --impl reference``) maps
none of this repo's native code into its process.
"""
import importlib

__version__ = "0.2.0"

_EXPORTS = {
    "MLP": "adapter", "MLP_Legacy": "adapter", "MLP_SD": "adapter", "MLP_plus": "adapter",
    "PEAAttn2Processor": "attn2", "attn2_core": "attn2", "install": "attn2",
    "kd_loss": "kd_loss", "BatchStager": "staging",
}
_SUBMODULES = ("_lib", "ops", "adapter", "attn2", "kd_loss", "staging", "train", "shell")


def __getattr__(name):
    if name in _EXPORTS:
        return getattr(importlib.import_module(f".{_EXPORTS[name]}", __name__), name)
    if name in _SUBMODULES:
        return importlib.import_module(f".{name}", __name__)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")


def __dir__():
    return sorted(list(globals()) + list(_EXPORTS) + list(_SUBMODULES))
