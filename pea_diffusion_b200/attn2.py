"""attn2 -- the frozen UNet's cross-attention blocks driven by the PEA context, as a diffusers attention processor.

Reference call path: ``train_sdxl_zh.py:397`` (student, autograd) and ``:415`` (teacher,
``no_grad``) -> diffusers ``Attention.forward`` -> ``AttnProcessor2_0.__call__``:
``q = to_q(h); k = to_k(ctx); v = to_v(ctx); o = SDPA(q, k, v); out = to_out[0](o)``.

``PEAAttn2Processor`` honours the same call protocol and is installed with ``unet.set_attn_processor`` (see
``install``).  What it replaces:
  * ``to_k`` / ``to_v``: ONE tcgen05 GEMM against the cached concatenation ``[Wk; Wv]`` -> K|V interleaved [B, L, 2C]
  * SDPA: the fused 77-key kernel (K/V resident in shared memory, Q/O streamed once, single-pass softmax)
  * backward: dQ and dK/dV from the fused backward kernel (S, P recomputed; no O / LSE re-read), then
    dContext = [dK | dV] @ [Wk; Wv] as one GEMM.  The projection weights are frozen: no weight gradient is ever
    formed (the reference computes and discards them -- SURVEY.md D6).
``to_q`` / ``to_out`` stay the module's own ``nn.Linear`` calls (cuBLAS; out of the hot-path scope, row K4).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import ops

Tensor = torch.Tensor
MAX_KEYS = 80
_HEAD_DIMS = (40, 48, 64, 80, 160)


class _Attn2Fn(torch.autograd.Function):
    """(q [B,N,C], ctx [B,L,Dc], Wkv [2C,Dc]) -> o [B,N,C]."""

    @staticmethod
    def forward(ctx_, q: Tensor, context: Tensor, wkv: Tensor, heads: int, scale: float, kv_cached: Optional[Tensor]):
        B, L, Dc = context.shape
        C2 = wkv.shape[0]
        if kv_cached is not None:
            kv = kv_cached
        else:
            kv = torch.empty(B, L, C2, device=q.device, dtype=q.dtype)
            ops.gemm([context.reshape(B * L, Dc)], [wkv], B * L, C2, Dc, out0=kv.view(B * L, C2))
        o = ops.attn2_fwd(q, kv, heads, scale)
        ctx_.save_for_backward(q, kv, wkv)
        ctx_.heads, ctx_.scale, ctx_.ctx_shape = heads, scale, (B, L, Dc)
        ctx_.kv_out = kv
        return o

    @staticmethod
    def backward(ctx_, do: Tensor):
        q, kv, wkv = ctx_.saved_tensors
        B, L, Dc = ctx_.ctx_shape
        dq, dkv = ops.attn2_bwd(q, kv, do, ctx_.heads, ctx_.scale)
        dctx = None
        if ctx_.needs_input_grad[1]:
            # dContext = dKV @ Wkv: A = dKV (two bf16/fp16 planes of the fp32 accumulators), B = Wkv stored [K=2C, N=Dc]
            planes = ops.split_planes(dkv.view(B * L, -1), 2, q.dtype)
            dctx = torch.empty(B, L, Dc, device=q.device, dtype=q.dtype)
            ops.gemm(planes, [wkv], B * L, Dc, wkv.shape[0], b_kmajor=False, out0=dctx.view(B * L, Dc))
        return (dq if ctx_.needs_input_grad[0] else None), dctx, None, None, None, None


def attn2_core(q: Tensor, context: Tensor, wkv: Tensor, heads: int, scale: Optional[float] = None,
               kv: Optional[Tensor] = None) -> Tensor:
    """Functional form: softmax(q (ctx Wk^T)^T * scale) (ctx Wv^T) with ``wkv = cat([Wk, Wv], 0)``."""
    d = q.shape[-1] // heads
    return _Attn2Fn.apply(q, context, wkv, heads, float(d ** -0.5 if scale is None else scale), kv)


class PEAAttn2Processor:
    """Drop-in for diffusers' ``AttnProcessor2_0`` on cross-attention (``*.attn2``) modules.

    Falls back to ``fallback`` (the stock processor) only for cases outside the kernel's contract -- self-attention,
    an attention mask, LoRA layers on ``to_k``/``to_v``, more than 80 keys, an unsupported head size, or non-16-bit
    activations.  CUDA is mandatory: host tensors raise (there is no CPU path).

    ``cache_kv=True`` keeps K|V per (context storage, version) so that a denoise loop that feeds the same context
    every step (tests/test_sdxl_zh.py:378-391) projects it once instead of once per step -- forward-only use.
    """

    def __init__(self, fallback=None, cache_kv: bool = False):
        self.fallback = fallback
        self.cache_kv = cache_kv
        self._wkv: Optional[Tensor] = None
        self._wkv_key = None
        self._kv: Optional[Tensor] = None
        self._kv_key = None

    # -- weights: [Wk; Wv] concatenated once per (weight version, dtype) ------------------------------------------
    def _weights(self, attn, dtype) -> Tensor:
        wk, wv = attn.to_k.weight, attn.to_v.weight
        key = (wk.data_ptr(), wv.data_ptr(), wk._version, wv._version, dtype, wk.device)
        if self._wkv is None or self._wkv_key != key:
            self._wkv = torch.cat([wk.detach(), wv.detach()], dim=0).to(dtype).contiguous()
            self._wkv_key = key
        return self._wkv

    def _supported(self, attn, hidden_states, encoder_hidden_states, attention_mask) -> bool:
        if encoder_hidden_states is None or attention_mask is not None:
            return False
        if hidden_states.dim() != 3 or hidden_states.dtype not in (torch.bfloat16, torch.float16):
            return False
        if getattr(attn.to_k, "lora_layer", None) is not None or getattr(attn.to_v, "lora_layer", None) is not None:
            return False
        if getattr(attn, "norm_cross", None) or getattr(attn, "group_norm", None) is not None or \
                getattr(attn, "spatial_norm", None) is not None or getattr(attn, "residual_connection", False) or \
                getattr(attn, "rescale_output_factor", 1.0) != 1.0:
            return False
        inner = attn.to_q.weight.shape[0]
        d = inner // attn.heads
        return encoder_hidden_states.shape[1] <= MAX_KEYS and d in _HEAD_DIMS and encoder_hidden_states.shape[-1] % 8 == 0 \
            and inner % 8 == 0

    def __call__(self, attn, hidden_states: Tensor, encoder_hidden_states: Optional[Tensor] = None,
                 attention_mask: Optional[Tensor] = None, temb: Optional[Tensor] = None, scale: float = 1.0) -> Tensor:
        if not self._supported(attn, hidden_states, encoder_hidden_states, attention_mask):
            if self.fallback is None:
                raise ops._lib.PeaError("PEAAttn2Processor: call outside the fused kernel's contract and no fallback processor set")
            return self.fallback(attn, hidden_states, encoder_hidden_states=encoder_hidden_states,
                                 attention_mask=attention_mask, temb=temb, scale=scale)
        ops.require_cuda(hidden_states, encoder_hidden_states)
        dt = hidden_states.dtype
        context = encoder_hidden_states.to(dt)
        if not context.is_contiguous():
            context = context.contiguous()
        q = attn.to_q(hidden_states)
        wkv = self._weights(attn, dt)
        kv = None
        if self.cache_kv and not (torch.is_grad_enabled() and context.requires_grad):
            key = (context.data_ptr(), context._version, tuple(context.shape), dt)
            if self._kv is None or self._kv_key != key:
                B, L, Dc = context.shape
                self._kv = torch.empty(B, L, wkv.shape[0], device=q.device, dtype=dt)
                ops.gemm([context.reshape(B * L, Dc)], [wkv], B * L, wkv.shape[0], Dc, out0=self._kv.view(B * L, -1))
                self._kv_key = key
            kv = self._kv
        o = _Attn2Fn.apply(q, context, wkv, attn.heads, float(attn.scale), kv)
        o = attn.to_out[0](o)
        return attn.to_out[1](o)


def install(unet, freeze: bool = True, cache_kv: bool = False) -> Dict[str, object]:
    """``unet.set_attn_processor({...})`` with ``PEAAttn2Processor`` on every ``*.attn2`` module and the existing
    processor kept everywhere else (diffusers requires the dict to cover every attention module).

    ``freeze=True`` also marks the UNet's parameters ``requires_grad_(False)``: the reference hands only the adapter's
    parameters to the optimizer (train_sdxl_zh.py:166-168) yet lets autograd form and discard every UNet weight
    gradient; skipping them does not change any result.
    """
    current = dict(unet.attn_processors)
    procs = {}
    for name, proc in current.items():
        procs[name] = PEAAttn2Processor(fallback=proc, cache_kv=cache_kv) if name.endswith("attn2.processor") else proc
    unet.set_attn_processor(procs)
    if freeze:
        unet.requires_grad_(False)
    return procs
