"""ctypes wrapper of ``oracle/pea_oracle_c.c`` -- the plain-C second oracle (TEST INFRASTRUCTURE ONLY).

``build()`` compiles it with gcc into ``oracle/_build/libpea_oracle_c.so`` (git-ignored; ``__graft_entry__.build()`` calls
it); the functions below take / return numpy float32 arrays.  Imported by tests/ only.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "pea_oracle_c.c")
LIB = os.path.join(_HERE, "_build", "libpea_oracle_c.so")
_lib: Optional[C.CDLL] = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.run(["gcc", "-O2", "-std=c99", "-fPIC", "-shared", "-fvisibility=hidden", "-o", LIB, SRC, "-lm"], check=True)
    return LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oc_adapter.restype = C.c_int
    return _lib


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def adapter(x, ln_w, ln_b, weights: Sequence[np.ndarray], fc_w=None, fc_b=None, use_residual: bool = False):
    """-> (x1 [B,out], x2 [B,L,out1]) when ``fc_w`` is given (SDXL), else x [B,L,out] (SD)."""
    x = _f32(x)
    B, L, in_dim = x.shape
    ws = [_f32(w) for w in weights]
    dims = (C.c_int * (len(ws) + 1))(in_dim, *[w.shape[0] for w in ws])
    wp = (C.c_void_p * len(ws))(*[w.ctypes.data for w in ws])
    out = ws[-1].shape[0]
    ln_w, ln_b = _f32(ln_w), _f32(ln_b)
    if fc_w is None:
        y = np.empty((B, L, out), np.float32)
        rc = lib().oc_adapter(_p(x), B, L, _p(ln_w), _p(ln_b), len(ws), wp, dims, None, None, 0, 0, None, _p(y))
        assert rc == 0
        return y
    fc_w, fc_b = _f32(fc_w), _f32(fc_b)
    x1 = np.empty((B, out), np.float32)
    x2 = np.empty((B, L, fc_w.shape[0]), np.float32)
    rc = lib().oc_adapter(_p(x), B, L, _p(ln_w), _p(ln_b), len(ws), wp, dims, _p(fc_w), _p(fc_b), fc_w.shape[0],
                          int(use_residual), _p(x1), _p(x2))
    assert rc == 0
    return x1, x2


def attn2_core(q, k, v, heads: int) -> np.ndarray:
    q, k, v = _f32(q), _f32(k), _f32(v)
    B, N, Cc = q.shape
    o = np.empty_like(q)
    lib().oc_attn2_core(_p(q), _p(k), _p(v), B, N, k.shape[1], heads, Cc // heads, _p(o))
    return o


def attn2_core_bwd(q, k, v, do, heads: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    q, k, v, do = _f32(q), _f32(k), _f32(v), _f32(do)
    B, N, Cc = q.shape
    dq, dk, dv = np.empty_like(q), np.empty_like(k), np.empty_like(v)
    lib().oc_attn2_core_bwd(_p(q), _p(k), _p(v), _p(do), B, N, k.shape[1], heads, Cc // heads, _p(dq), _p(dk), _p(dv))
    return dq, dk, dv


def kd_loss(pred, noise, pred_t, fs: Sequence[np.ndarray], ft: Sequence[np.ndarray], zh, feature_weight: float = 0.1,
            skip_nonfinite: bool = False):
    """-> ((L, L0, L1, L2), dpred, [d f_S])"""
    pred, noise, pred_t = _f32(pred), _f32(noise), _f32(pred_t)
    fs, ft = [_f32(a) for a in fs], [_f32(a) for a in ft]
    B = pred.shape[0]
    zh = np.ascontiguousarray(np.asarray(zh, dtype=np.int64))
    n = len(fs)
    numel = (C.c_longlong * max(n, 1))(*[a.size // B for a in fs])
    fsp = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in fs])
    ftp = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in ft])
    dfs: List[np.ndarray] = [np.empty_like(a) for a in fs]
    dfp = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in dfs])
    dpred = np.empty_like(pred)
    out = (C.c_double * 4)()
    lib().oc_kd_loss(_p(pred), _p(noise), _p(pred_t), B, C.c_longlong(pred.size // B), n, fsp, ftp, numel,
                     zh.ctypes.data_as(C.c_void_p), C.c_double(feature_weight), int(skip_nonfinite), out, _p(dpred), dfp)
    return tuple(out), dpred, dfs
