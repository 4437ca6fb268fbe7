"""Thin, allocation-aware Python wrappers over the C ABI (one function per entry point of ``include/pea_b200.h``).

These are the only places that call into ``libpea_b200.so``; the autograd Functions in ``adapter.py``, ``attn2.py``
and ``kd_loss.py`` are built from them.  Nothing here falls back to PyTorch math.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import GemmDesc, Attn2Desc, KdTap, check, dtype_code, lib, planes3, ptr, require_cuda, stream_ptr

Tensor = torch.Tensor
Planes = List[Tensor]


class KernelTimer:
    """Optional per-launch CUDA-event timing of the library's kernels on the launching stream (used by bench.py for the
    roofline figures).  Disabled by default: zero overhead on the product path."""

    def __init__(self):
        self.enabled = False
        self.external = False
        self.records = []  # (name, algorithmic_bytes, flops, start_event, end_event)

    def reset(self, enabled: bool, external: bool = False) -> None:
        """``external=True``: the events become event-record NODES when the launches are captured into a CUDA graph, so
        every replay of that graph re-records them and the kernels can be timed inside the replayed step."""
        self.enabled = enabled
        self.external = external
        self.records = []

    def summary(self):
        """-> {name: {"launches", "ms", "bytes", "flops"}}; call after a device synchronize."""
        out = {}
        for name, nbytes, flops, e0, e1 in self.records:
            r = out.setdefault(name, {"launches": 0, "ms": 0.0, "bytes": 0, "flops": 0})
            r["launches"] += 1
            r["ms"] += e0.elapsed_time(e1)
            r["bytes"] += nbytes
            r["flops"] += flops
        return out


TIMER = KernelTimer()


class _timed:
    def __init__(self, name: str, nbytes: int, flops: int = 0):
        self.args = (name, nbytes, flops)

    def __enter__(self):
        if TIMER.enabled:
            self.e0 = torch.cuda.Event(enable_timing=True, external=True) if TIMER.external else torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if TIMER.enabled:
            e1 = torch.cuda.Event(enable_timing=True, external=True) if TIMER.external else torch.cuda.Event(enable_timing=True)
            e1.record()
            TIMER.records.append((*self.args, self.e0, e1))
        return False


def _row_major_2d(t: Tensor) -> None:
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"expected a 2-D tensor with unit column stride, got shape {tuple(t.shape)} strides {t.stride()}")


def gemm(A: Sequence[Tensor], B: Sequence[Tensor], M: int, N: int, K: int, *, a_kmajor: bool = True, b_kmajor: bool = True,
         alpha: float = 1.0, bias: Optional[Tensor] = None, aux: Optional[Tensor] = None,
         rowadd: Optional[Tensor] = None, rowadd_group: int = 1, rowadd_scale: float = 1.0, act: bool = False,
         out0: Optional[Tensor] = None, accumulate: bool = False, out1: Optional[Sequence[Tensor]] = None,
         split_k: int = 0, pdl: bool = False) -> None:
    """C[m,n] = sum_k A(m,k) B(n,k) on tcgen05 tensor cores; see ``pea_gemm`` in the header for the epilogue.

    ``A`` / ``B`` are lists of 1..3 16-bit planes (2-D, unit column stride).  K-major A is stored [M, K]; MN-major A is
    stored [K, M] (i.e. pass the un-transposed row-major tensor of A^T).  Same for B with N.
    """
    require_cuda(*A, *B, out0, bias, aux, rowadd, *(out1 or []))
    for t in list(A) + list(B):
        _row_major_2d(t)
    d = GemmDesc()
    d.M, d.N, d.K = M, N, K
    d.in_dtype = dtype_code(A[0].dtype)
    if any(t.dtype != A[0].dtype for t in list(A) + list(B)):
        raise TypeError("gemm: all operand planes must share one 16-bit dtype")
    d.nsplit_a, d.nsplit_b = len(A), len(B)
    d.d_A = planes3(A)
    d.d_B = planes3(B)
    d.a_kmajor, d.b_kmajor = int(a_kmajor), int(b_kmajor)
    d.lda, d.ldb = A[0].stride(0), B[0].stride(0)
    if any(t.stride(0) != d.lda for t in A) or any(t.stride(0) != d.ldb for t in B):
        raise ValueError("gemm: planes of one operand must share a row stride")
    d.alpha = alpha
    if bias is not None:
        if bias.dtype != torch.float32 or not bias.is_contiguous():
            raise TypeError("gemm: bias must be contiguous fp32")
        d.d_bias = bias.data_ptr()
    if aux is not None:
        _row_major_2d(aux)
        d.d_aux, d.aux_dtype, d.ldaux = aux.data_ptr(), dtype_code(aux.dtype), aux.stride(0)
    if rowadd is not None:
        if rowadd.dtype != torch.float32 or not rowadd.is_contiguous() or rowadd.shape[-1] != N:
            raise TypeError("gemm: rowadd must be contiguous fp32 [*, N]")
        d.d_rowadd, d.rowadd_group, d.rowadd_scale = rowadd.data_ptr(), rowadd_group, rowadd_scale
    d.act = int(act)
    if out0 is not None:
        _row_major_2d(out0)
        d.d_out0, d.out0_dtype, d.ldo0, d.out0_accumulate = out0.data_ptr(), dtype_code(out0.dtype), out0.stride(0), int(accumulate)
    if out1:
        for t in out1:
            _row_major_2d(t)
        d.d_out1 = planes3(out1)
        d.out1_nsplit, d.out1_dtype, d.ldo1 = len(out1), dtype_code(out1[0].dtype), out1[0].stride(0)
    d.split_k = split_k
    d.flags = _lib.PEA_LAUNCH_PDL if pdl else 0
    nprod = sum(1 for i in range(len(A)) for j in range(len(B)) if i + j < max(len(A), len(B)))
    io = 2 * (len(A) * M * K + len(B) * N * K) + (out0.element_size() * M * N if out0 is not None else 0) + \
        (sum(t.element_size() for t in out1) * M * N if out1 else 0) + (aux.element_size() * M * N if aux is not None else 0)
    with _timed("gemm", io, 2 * M * N * K * nprod):
        check(lib.pea_gemm(C.byref(d), stream_ptr(A[0].device)), "pea_gemm")


def _flags(pdl: bool) -> int:
    return _lib.PEA_LAUNCH_PDL if pdl else 0


def split_planes(x: Tensor, nsplit: int, plane_dtype: torch.dtype, pdl: bool = False) -> Planes:
    """x [rows, cols] (any float dtype, unit column stride) -> ``nsplit`` dense 16-bit planes summing to x."""
    require_cuda(x)
    _row_major_2d(x)
    rows, cols = x.shape
    if nsplit == 1 and x.dtype == plane_dtype and x.stride(0) % 8 == 0 and x.data_ptr() % 16 == 0:
        return [x]  # already a valid single plane: no copy
    planes = [torch.empty(rows, cols, device=x.device, dtype=plane_dtype) for _ in range(nsplit)]
    check(lib.pea_split_planes(x.data_ptr(), dtype_code(x.dtype), rows, cols, x.stride(0), C.byref(planes3(planes)), nsplit,
                               dtype_code(plane_dtype), _flags(pdl), stream_ptr(x.device)), "pea_split_planes")
    return planes


def split_planes_multi(xs: Sequence[Tensor], nsplit: int, plane_dtype: torch.dtype, pdl: bool = False) -> List[Planes]:
    """split_planes of several tensors of one dtype -- ONE launch when all are dense fp32 (the adapter's weights, re-split
    every step)."""
    require_cuda(*xs)
    out: List[Optional[Planes]] = [None] * len(xs)
    todo = []
    for i, x in enumerate(xs):
        _row_major_2d(x)
        if nsplit == 1 and x.dtype == plane_dtype and x.stride(0) % 8 == 0 and x.data_ptr() % 16 == 0:
            out[i] = [x]
        else:
            todo.append(i)
    by_dtype = {}
    for i in todo:
        by_dtype.setdefault(xs[i].dtype, []).append(i)
    for dt, idx in by_dtype.items():
        for k in range(0, len(idx), 8):
            chunk = idx[k:k + 8]
            items = (_lib.SplitItem * len(chunk))()
            for it, i in zip(items, chunk):
                x = xs[i]
                out[i] = [torch.empty(x.shape[0], x.shape[1], device=x.device, dtype=plane_dtype) for _ in range(nsplit)]
                it.d_x, it.rows, it.cols, it.ld = x.data_ptr(), x.shape[0], x.shape[1], x.stride(0)
                for s, pl in enumerate(out[i]):
                    it.d_planes[s] = pl.data_ptr()
            check(lib.pea_split_planes_multi(C.cast(items, C.c_void_p), len(chunk), dtype_code(dt), nsplit, dtype_code(plane_dtype),
                                             _flags(pdl), stream_ptr(xs[chunk[0]].device)), "pea_split_planes_multi")
    return out


def layernorm_fwd(x: Tensor, gamma: Tensor, beta: Tensor, eps: float, nsplit: int, plane_dtype: torch.dtype, pdl: bool = False):
    """x [M, D] contiguous -> (planes of LN(x), mean [M], rstd [M]); statistics in fp32."""
    require_cuda(x, gamma, beta)
    M, D = x.shape
    planes = [torch.empty(M, D, device=x.device, dtype=plane_dtype) for _ in range(nsplit)]
    mean = torch.empty(M, device=x.device, dtype=torch.float32)
    rstd = torch.empty(M, device=x.device, dtype=torch.float32)
    check(lib.pea_layernorm_fwd(x.data_ptr(), dtype_code(x.dtype), gamma.data_ptr(), beta.data_ptr(), M, D, eps,
                                C.byref(planes3(planes)), nsplit, dtype_code(plane_dtype), mean.data_ptr(), rstd.data_ptr(),
                                _flags(pdl), stream_ptr(x.device)), "pea_layernorm_fwd")
    return planes, mean, rstd


def layernorm_bwd(x: Tensor, dy: Tensor, gamma: Tensor, mean: Tensor, rstd: Tensor, dgamma: Tensor, dbeta: Tensor,
                  want_dx: bool, pdl: bool = False) -> Optional[Tensor]:
    """dgamma/dbeta (fp32) are accumulated into; returns dx (fp32) when asked."""
    require_cuda(x, dy)
    M, D = x.shape
    dx = torch.empty(M, D, device=x.device, dtype=torch.float32) if want_dx else None
    check(lib.pea_layernorm_bwd(x.data_ptr(), dtype_code(x.dtype), dy.data_ptr(), gamma.data_ptr(), mean.data_ptr(),
                                rstd.data_ptr(), M, D, dgamma.data_ptr(), dbeta.data_ptr(), ptr(dx), _flags(pdl),
                                stream_ptr(x.device)),
          "pea_layernorm_bwd")
    return dx


def mean_tokens(x: Tensor, B: int, L: int, res: Optional[Tensor] = None, pdl: bool = False) -> Tensor:
    """x fp32 [B*L, D] (+ res [B*L, D]) -> [B, D] fp32 mean over the L tokens of each sample."""
    require_cuda(x, res)
    D = x.shape[-1]
    out = torch.empty(B, D, device=x.device, dtype=torch.float32)
    check(lib.pea_mean_tokens(x.data_ptr(), ptr(res), dtype_code(res.dtype) if res is not None else 0, B, L, D,
                              out.data_ptr(), _flags(pdl), stream_ptr(x.device)), "pea_mean_tokens")
    return out


def colsum_accum(x: Tensor, out: Tensor, pdl: bool = False) -> None:
    """out[n] += sum_m x[m, n]; out fp32."""
    require_cuda(x, out)
    _row_major_2d(x)
    check(lib.pea_colsum_accum(x.data_ptr(), dtype_code(x.dtype), x.shape[0], x.shape[1], x.stride(0), out.data_ptr(),
                               _flags(pdl), stream_ptr(x.device)), "pea_colsum_accum")


def cfg_select(mask_u8: Tensor, uncond: Tensor, cond: Tensor) -> Tensor:
    """out[b] = mask[b] ? uncond[b] : cond[b]  (train_sdxl_zh.py:392-395)."""
    require_cuda(mask_u8, uncond, cond)
    if uncond.shape != cond.shape or uncond.dtype != cond.dtype:
        raise ValueError("cfg_select: cond / uncond must match in shape and dtype")
    uncond, cond = uncond.contiguous(), cond.contiguous()
    out = torch.empty_like(cond)
    B = cond.shape[0]
    check(lib.pea_cfg_select(mask_u8.data_ptr(), uncond.data_ptr(), cond.data_ptr(), out.data_ptr(), B, cond.numel() // B,
                             dtype_code(cond.dtype), stream_ptr(cond.device)), "pea_cfg_select")
    return out


def add_noise(latents: Tensor, noise_raw: Tensor, offs: Optional[Tensor], noise_offset: float, alphas_cumprod: Tensor,
              timesteps: Tensor, out_dtype: torch.dtype):
    """Offset noise + DDPM add_noise + cast (train_sdxl_zh.py:311-323) -> (noise, noisy_latents) in ``out_dtype``."""
    require_cuda(latents, noise_raw, offs, alphas_cumprod, timesteps)
    B, Cc = latents.shape[:2]
    HW = latents.numel() // (B * Cc)
    latents = latents.contiguous()
    noise_raw = noise_raw.to(torch.float32).contiguous()
    offs_c = None if offs is None else offs.to(torch.float32).reshape(B, Cc).contiguous()
    acp = alphas_cumprod.to(torch.float32).contiguous()
    t = timesteps.to(torch.int64).contiguous()
    noise = torch.empty(latents.shape, device=latents.device, dtype=out_dtype)
    noisy = torch.empty(latents.shape, device=latents.device, dtype=out_dtype)
    check(lib.pea_add_noise(latents.data_ptr(), dtype_code(latents.dtype), noise_raw.data_ptr(), ptr(offs_c), float(noise_offset),
                            acp.data_ptr(), acp.numel(), t.data_ptr(), B, Cc, HW, noise.data_ptr(), noisy.data_ptr(),
                            dtype_code(out_dtype), stream_ptr(latents.device)), "pea_add_noise")
    return noise, noisy


def attn2_fwd(q: Tensor, kv: Tensor, heads: int, scale: float, pdl: bool = False) -> Tensor:
    """q [B,N,C], kv [B,L,2C] (K in columns [0,C), V in [C,2C)) -> o [B,N,C] = softmax(q k^T scale) v per head."""
    require_cuda(q, kv)
    B, N, Cq = q.shape
    L = kv.shape[1]
    o = torch.empty(B, N, Cq, device=q.device, dtype=q.dtype)
    d = _attn_desc(q, kv, heads, scale)
    d.flags = _lib.PEA_LAUNCH_PDL if pdl else 0
    d.d_o, d.ld_o = o.data_ptr(), Cq
    with _timed("attn2_fwd", attn2_bytes(B, N, L, Cq, q.element_size())[0], 4 * B * N * L * Cq):
        check(lib.pea_attn2_fwd(C.byref(d), stream_ptr(q.device)), "pea_attn2_fwd")
    return o


def attn2_bwd(q: Tensor, kv: Tensor, do: Tensor, heads: int, scale: float, dkv: Optional[Tensor] = None, pdl: bool = False):
    """-> (dq [B,N,C] in q's dtype, dkv [B,L,2C] fp32).  ``dkv`` may be a caller-owned fp32 [B,L,2C] view (unit column
    stride) that is ACCUMULATED into -- e.g. one layer's column slice of a buffer shared by all layers."""
    require_cuda(q, kv, do, dkv)
    B, N, Cq = q.shape
    L = kv.shape[1]
    if do.stride(-1) != 1 or do.stride(1) % 8 != 0 or do.stride(0) != N * do.stride(1):
        do = do.contiguous()
    dq = torch.empty(B, N, Cq, device=q.device, dtype=q.dtype)
    if dkv is None:
        dkv = torch.zeros(B, L, 2 * Cq, device=q.device, dtype=torch.float32)
    elif dkv.dtype != torch.float32 or tuple(dkv.shape) != (B, L, 2 * Cq) or dkv.stride(2) != 1:
        raise ValueError("attn2_bwd: dkv must be an fp32 [B, L, 2C] view with unit column stride")
    d = _attn_desc(q, kv, heads, scale)
    d.flags = _lib.PEA_LAUNCH_PDL if pdl else 0
    d.d_do, d.ld_do = do.data_ptr(), do.stride(1)
    d.d_dq, d.ld_dq = dq.data_ptr(), Cq
    d.d_dk, d.d_dv = dkv.data_ptr(), dkv.data_ptr() + 4 * Cq
    d.ld_dkv, d.bs_dkv = dkv.stride(1), dkv.stride(0)
    with _timed("attn2_bwd", attn2_bytes(B, N, L, Cq, q.element_size())[1], 10 * B * N * L * Cq):
        check(lib.pea_attn2_bwd(C.byref(d), stream_ptr(q.device)), "pea_attn2_bwd")
    return dq, dkv


def attn2_bytes(B: int, N: int, L: int, C: int, es: int = 2):
    """Algorithmic HBM bytes of one attn2 core call, SURVEY.md section 8d: forward = read Q, write O, read K, V;
    backward = read Q, dO, write dQ, read K, V, write dK, dV.  (dK / dV are really accumulated in fp32 -- 8 bytes per
    element instead of the formula's 2 x e -- which the roofline figures do NOT credit.)  -> (fwd, bwd)."""
    return 2 * B * N * C * es + 2 * B * L * C * es, 3 * B * N * C * es + 4 * B * L * C * es


def _attn_desc(q: Tensor, kv: Tensor, heads: int, scale: float) -> Attn2Desc:
    B, N, Cq = q.shape
    if kv.dim() != 3 or kv.shape[0] != B or kv.shape[2] != 2 * Cq or kv.dtype != q.dtype:
        raise ValueError(f"attn2: kv must be [B, L, 2C] in q's dtype; q {tuple(q.shape)} {q.dtype}, kv {tuple(kv.shape)} {kv.dtype}")
    if q.stride(2) != 1 or q.stride(0) != N * q.stride(1) or kv.stride(2) != 1:
        raise ValueError("attn2: q must be [B,N,C] with unit column stride and dense batch stride; kv unit column stride")
    d = Attn2Desc()
    d.B, d.N, d.L, d.H, d.d = B, N, kv.shape[1], heads, Cq // heads
    d.scale, d.dtype = scale, dtype_code(q.dtype)
    d.d_q, d.ld_q = q.data_ptr(), q.stride(1)
    es = q.element_size()
    d.d_k, d.d_v = kv.data_ptr(), kv.data_ptr() + Cq * es
    d.ld_kv, d.bs_kv = kv.stride(1), kv.stride(0)
    return d


class KdTaps:
    """Host-side array of ``pea_kd_tap`` records plus the tensors they point into (kept alive)."""

    def __init__(self):
        self.arr = (KdTap * _lib.KD_MAX_TAPS)()
        self.n = 0
        self.keep: List[Tensor] = []

    def add(self, student: Tensor, teacher: Tensor, weight: Tensor, group: int, grad: Optional[Tensor] = None,
            teacher2: Optional[Tensor] = None, weight2: Optional[Tensor] = None, group2: int = 0) -> None:
        if self.n >= _lib.KD_MAX_TAPS:
            raise ValueError(f"at most {_lib.KD_MAX_TAPS} KD taps per call")
        t = self.arr[self.n]
        t.d_student, t.d_teacher, t.d_teacher2 = student.data_ptr(), teacher.data_ptr(), ptr(teacher2)
        t.d_grad, t.d_weight, t.d_weight2 = ptr(grad), weight.data_ptr(), ptr(weight2)
        t.numel_per_sample = student.numel() // student.shape[0]
        t.group, t.group2 = group, group2
        self.keep += [x for x in (student, teacher, teacher2, grad, weight, weight2) if x is not None]
        self.n += 1


def kd_mse_fwd_bwd(taps: KdTaps, B: int, dtype: torch.dtype, feature_weight: float, grad_scale: float, skip_nonfinite: bool,
                   device: torch.device) -> Tensor:
    """-> out fp32 [4 + 2*ntaps] (layout in the header).  Gradients land in the taps' ``grad`` tensors."""
    ws_bytes = lib.pea_kd_workspace_bytes(taps.arr, taps.n, B)
    ws = torch.empty(max(ws_bytes // 4, 1), device=device, dtype=torch.float32)
    out = torch.empty(4 + 2 * taps.n, device=device, dtype=torch.float32)
    es = 4 if dtype == torch.float32 else 2
    nbytes = 0
    for i in range(taps.n):
        t = taps.arr[i]
        nbytes += t.numel_per_sample * B * es * (2 + (1 if t.d_teacher2 else 0) + (1 if t.d_grad else 0))
    with _timed("kd_mse", nbytes):
        check(lib.pea_kd_mse_fwd_bwd(taps.arr, taps.n, B, dtype_code(dtype), feature_weight, grad_scale, int(skip_nonfinite),
                                     ws.data_ptr(), out.data_ptr(), stream_ptr(device)), "pea_kd_mse_fwd_bwd")
    return out


def kd_scale_grads(taps: KdTaps, B: int, dtype: torch.dtype, upstream: Optional[Tensor], out: Tensor, device: torch.device,
                   inv_prescale: float = 1.0) -> None:
    check(lib.pea_kd_scale_grads(taps.arr, taps.n, B, dtype_code(dtype), ptr(upstream), float(inv_prescale), out.data_ptr(),
                                 stream_ptr(device)), "pea_kd_scale_grads")
