"""ctypes binding of ``libpea_b200.so`` (the C ABI declared in ``include/pea_b200.h``).

There is no CPU or PyTorch fallback behind these calls: if the library has not been built
(``python -c "import __graft_entry__ as g; g.build()"``) importing this module raises, and every compute entry point
fails with the library's own error text when no sm_100 device is present.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PEA_B200_LIB") or os.path.join(_HERE, "libpea_b200.so")   # override: instrumented debug builds

PEA_F32, PEA_F16, PEA_BF16 = 0, 1, 2
PEA_LAUNCH_PDL = 1
KD_MAX_TAPS = 16

_DTYPE = {torch.float32: PEA_F32, torch.float16: PEA_F16, torch.bfloat16: PEA_BF16}


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _DTYPE[dt]
    except KeyError:
        raise TypeError(f"libpea_b200 supports float32/float16/bfloat16 tensors, got {dt}") from None


class PeaError(RuntimeError):
    pass


class KdTap(C.Structure):
    _fields_ = [
        ("d_student", C.c_void_p), ("d_teacher", C.c_void_p), ("d_teacher2", C.c_void_p), ("d_grad", C.c_void_p),
        ("d_weight", C.c_void_p), ("d_weight2", C.c_void_p), ("numel_per_sample", C.c_longlong),
        ("group", C.c_int), ("group2", C.c_int),
    ]


class SplitItem(C.Structure):
    _fields_ = [("d_x", C.c_void_p), ("rows", C.c_longlong), ("cols", C.c_longlong), ("ld", C.c_longlong), ("d_planes", C.c_void_p * 3)]


class Attn2Desc(C.Structure):
    _fields_ = [
        ("B", C.c_int), ("N", C.c_int), ("L", C.c_int), ("H", C.c_int), ("d", C.c_int),
        ("scale", C.c_float), ("dtype", C.c_int),
        ("d_q", C.c_void_p), ("ld_q", C.c_longlong),
        ("d_k", C.c_void_p), ("d_v", C.c_void_p), ("ld_kv", C.c_longlong), ("bs_kv", C.c_longlong),
        ("d_o", C.c_void_p), ("ld_o", C.c_longlong),
        ("d_do", C.c_void_p), ("ld_do", C.c_longlong),
        ("d_dq", C.c_void_p), ("ld_dq", C.c_longlong),
        ("d_dk", C.c_void_p), ("d_dv", C.c_void_p), ("ld_dkv", C.c_longlong), ("bs_dkv", C.c_longlong),
        ("flags", C.c_uint),
    ]


class GemmDesc(C.Structure):
    _fields_ = [
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("in_dtype", C.c_int), ("nsplit_a", C.c_int), ("nsplit_b", C.c_int),
        ("d_A", C.c_void_p * 3), ("a_kmajor", C.c_int), ("lda", C.c_longlong),
        ("d_B", C.c_void_p * 3), ("b_kmajor", C.c_int), ("ldb", C.c_longlong),
        ("alpha", C.c_float),
        ("d_bias", C.c_void_p),
        ("d_aux", C.c_void_p), ("aux_dtype", C.c_int), ("ldaux", C.c_longlong),
        ("d_rowadd", C.c_void_p), ("rowadd_group", C.c_int), ("rowadd_scale", C.c_float),
        ("act", C.c_int),
        ("d_out0", C.c_void_p), ("out0_dtype", C.c_int), ("ldo0", C.c_longlong), ("out0_accumulate", C.c_int),
        ("d_out1", C.c_void_p * 3), ("out1_nsplit", C.c_int), ("out1_dtype", C.c_int), ("ldo1", C.c_longlong),
        ("split_k", C.c_int),
        ("flags", C.c_uint),
    ]


# every symbol include/pea_b200.h declares: name -> (restype, argtypes)
_VP3 = C.c_void_p * 3
SYMBOLS = {
    "pea_version": (C.c_int, []),
    "pea_last_error": (C.c_char_p, []),
    "pea_launch_count": (C.c_longlong, []),
    "pea_kd_workspace_bytes": (C.c_size_t, [C.POINTER(KdTap), C.c_int, C.c_int]),
    "pea_kd_mse_fwd_bwd": (C.c_int, [C.POINTER(KdTap), C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_void_p]),
    "pea_kd_scale_grads": (C.c_int, [C.POINTER(KdTap), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]),
    "pea_attn2_fwd": (C.c_int, [C.POINTER(Attn2Desc), C.c_void_p]),
    "pea_attn2_bwd": (C.c_int, [C.POINTER(Attn2Desc), C.c_void_p]),
    "pea_gemm": (C.c_int, [C.POINTER(GemmDesc), C.c_void_p]),
    "pea_layernorm_fwd": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float,
                                    C.POINTER(_VP3), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p]),
    "pea_layernorm_bwd": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p]),
    "pea_split_planes": (C.c_int, [C.c_void_p, C.c_int, C.c_longlong, C.c_longlong, C.c_longlong, C.POINTER(_VP3), C.c_int,
                                   C.c_int, C.c_uint, C.c_void_p]),
    "pea_split_planes_multi": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_void_p]),
    "pea_mean_tokens": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_uint, C.c_void_p]),
    "pea_colsum_accum": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_void_p, C.c_uint, C.c_void_p]),
    "pea_add_noise": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                C.c_int, C.c_longlong, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "pea_cfg_select": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_int, C.c_void_p]),
}


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA library first (python -c 'import __graft_entry__ as g; g.build()'). "
            "pea_diffusion_b200 has no CPU / PyTorch fallback for its kernels.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise PeaError(f"{what} failed (code {rc}): {lib.pea_last_error().decode(errors='replace')}")


def stream_ptr(device: Optional[torch.device] = None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors: torch.Tensor) -> None:
    """Every operand on the CURRENT CUDA device: the C side sizes grids, sets function attributes and launches on the
    process's current device (one process per GPU), so a tensor on another device would run on a foreign stream."""
    cur = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise PeaError("pea_diffusion_b200 kernels need CUDA tensors (there is no CPU fallback); got a "
                           f"{t.device} tensor of shape {tuple(t.shape)}")
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise PeaError(f"tensor on {t.device} but the current CUDA device is cuda:{cur}; one process per GPU -- call "
                           "torch.cuda.set_device(...) (or wrap the call in torch.cuda.device(...)) first")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def planes3(ts: Sequence[torch.Tensor]) -> "_VP3":
    arr = _VP3()
    for i, t in enumerate(ts):
        arr[i] = t.data_ptr()
    return arr


def launch_count() -> int:
    return int(lib.pea_launch_count())
