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
``to_q`` / ``to_out[0]`` run on the same tcgen05 GEMM with the bias in its epilogue, and the three
kernels of a layer -- to_q GEMM, attention core, to_out GEMM -- are launched as programmatic dependents of each other
(PEA_LAUNCH_PDL) so the attention kernel's prologue hides under its own producer's tail; ``own_projections=False`` keeps
the module's ``nn.Linear`` calls (cuBLAS).  The residual add is outside the processor in diffusers' block, so it stays there.
"""
from __future__ import annotations

import warnings
import weakref
from typing import Dict, Optional

import torch

from . import ops

Tensor = torch.Tensor
MAX_KEYS = 80
_HEAD_DIMS = (40, 48, 64, 80, 160)
PAD_SMALL_HEADS = True     # head_dim 40 / 48 (SD1.5's 320-channel blocks) run on the tcgen05 kernels, zero-padded to 64


def _padded_head_dim(d: int) -> int:
    """Head sizes below 64 are zero-padded to 64 IN THE FROZEN PROJECTION WEIGHTS (to_q / to_k / to_v get zero output rows,
    to_out zero input columns), so q / k / v / o simply come out 64 wide per head and the head_dim-64 tcgen05 kernels serve
    them: padded dims add 0 to q.k, produce 0 in o and meet zero columns of to_out.  The softmax scale stays d^-1/2."""
    return 64 if (PAD_SMALL_HEADS and d < 64 and d % 8 == 0) else d


def _pad_rows(w: Tensor, heads: int, d: int, dp: int) -> Tensor:
    """[H*d, X] -> [H*dp, X]: every head's block of d output rows followed by dp - d zero rows."""
    if dp == d:
        return w
    out = w.new_zeros(heads, dp, w.shape[1])
    out[:, :d] = w.view(heads, d, w.shape[1])
    return out.view(heads * dp, w.shape[1])


def _pad_cols(w: Tensor, heads: int, d: int, dp: int) -> Tensor:
    """[X, H*d] -> [X, H*dp]: zero input columns for the padded head dims."""
    if dp == d:
        return w
    out = w.new_zeros(w.shape[0], heads, dp)
    out[:, :, :d] = w.view(w.shape[0], heads, d)
    return out.view(w.shape[0], heads * dp)


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


class _FrozenLinearFn(torch.autograd.Function):
    """y = x W^T (+ b) for a FROZEN ``nn.Linear`` (``attn.to_q`` / ``attn.to_out[0]``, SURVEY.md kernel K4) on the library's
    tcgen05 GEMM: bias in the epilogue, dx = dy W in the backward, no weight / bias gradient (the reference forms and
    discards them).  Launched with PEA_LAUNCH_PDL so that  to_q -> attn2 core -> to_out  (and the mirrored
    chain of the backward) runs as programmatic dependents: each kernel's prologue overlaps its predecessor's tail."""

    @staticmethod
    def forward(ctx_, x: Tensor, w: Tensor, b32: Optional[Tensor]):
        shp = x.shape
        x2 = x.reshape(-1, shp[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        y = torch.empty(x2.shape[0], w.shape[0], device=x.device, dtype=x.dtype)
        ops.gemm([x2], [w], x2.shape[0], w.shape[0], w.shape[1], bias=b32, out0=y, pdl=True)
        ctx_.save_for_backward(w)
        ctx_.shape = shp
        return y.view(*shp[:-1], w.shape[0])

    @staticmethod
    def backward(ctx_, dy: Tensor):
        (w,) = ctx_.saved_tensors
        dy2 = dy.reshape(-1, w.shape[0])
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        dx = torch.empty(dy2.shape[0], w.shape[1], device=dy.device, dtype=dy.dtype)
        ops.gemm([dy2], [w], dy2.shape[0], w.shape[1], w.shape[0], b_kmajor=False, out0=dx, pdl=True)
        return dx.view(ctx_.shape), None, None


class _ProjectAllFn(torch.autograd.Function):
    """context [B,L,Dc] -> (K|V of EVERY attn2 layer of the UNet as one [B*L, sum(2C)] tensor, gradient token).

    One GEMM against the concatenation of all layers' [Wk; Wv] replaces 70 (SDXL) one-wave GEMMs per UNet forward, and
    its backward forms dContext = dKV_all @ Wcat_all as ONE split-K GEMM instead of 70 GEMMs + 69 accumulations.  The
    per-layer attention backward kernels add their dK|dV straight into column slices of ``group.dkv_all``; the 1-element
    ``token`` output is consumed only by the FIRST attn2 layer of the forward pass, whose backward is the last attn2
    backward to run (every other attn2 layer is downstream of it), so autograd calls this node's backward exactly when
    ``dkv_all`` is complete."""

    @staticmethod
    def forward(ctx_, context: Tensor, group: "Attn2Group"):
        B, L, Dc = context.shape
        wcat = group.weights(context.dtype, context.device)
        kv_all = torch.empty(B * L, group.total, device=context.device, dtype=context.dtype)
        ops.gemm([context.reshape(B * L, Dc)], [wcat], B * L, group.total, Dc, out0=kv_all)
        ctx_.group, ctx_.shape, ctx_.dtype = group, (B, L, Dc), context.dtype
        # per-pass state (several passes may be outstanding before their backwards run, e.g. gradient accumulation)
        state = _PassState()
        if ctx_.needs_input_grad[0]:
            state.dkv_all = torch.zeros(B, L, group.total, device=context.device, dtype=torch.float32)
        ctx_.state = group.state = state
        token = context.new_zeros(1)
        ctx_.mark_non_differentiable(kv_all)
        return kv_all, token

    @staticmethod
    def backward(ctx_, _g_kv, g_token):
        group, state = ctx_.group, ctx_.state
        B, L, Dc = ctx_.shape
        if state.bwd_calls != state.fwd_calls:
            raise RuntimeError(f"attn2 group: dContext requested after {state.bwd_calls} of {state.fwd_calls} layer "
                               "backwards -- the first attn2 layer is not upstream of all others in this network")
        wcat = group.weights(ctx_.dtype, state.dkv_all.device)
        # two 16-bit planes (hi + residual) of the fp32 accumulators, exactly as the per-layer path does: rounding the sum
        # over all layers' query tiles to ONE bf16 plane before the dContext GEMM costs ~3 decimal digits for nothing
        dkv16 = ops.split_planes(state.dkv_all.view(B * L, group.total), group.dctx_planes, ctx_.dtype)
        dctx = torch.zeros(B * L, Dc, device=wcat.device, dtype=torch.float32)
        kb = (group.total + 63) // 64
        tiles = ((B * L + 127) // 128) * ((Dc + 127) // 128)
        split = max(1, min(kb // 8, (4 * group.sms + tiles - 1) // tiles))
        ops.gemm(dkv16, [wcat], B * L, Dc, group.total, b_kmajor=False, out0=dctx, accumulate=True, split_k=split)
        state.dkv_all = None
        return dctx.view(B, L, Dc).to(ctx_.dtype), None


class _Attn2GroupedFn(torch.autograd.Function):
    """One layer's attention core on its column slice of the group's K|V tensor."""

    @staticmethod
    def forward(ctx_, q: Tensor, token: Optional[Tensor], kv: Tensor, group: "Attn2Group", layer: int, heads: int, scale: float,
                pdl: bool = False):
        o = ops.attn2_fwd(q, kv, heads, scale, pdl=pdl)
        ctx_.save_for_backward(q, kv)
        ctx_.group, ctx_.layer, ctx_.heads, ctx_.scale, ctx_.has_token = group, layer, heads, scale, token is not None
        ctx_.pdl = pdl
        ctx_.state = group.state
        ctx_.track = ctx_.state is not None and ctx_.state.dkv_all is not None
        if ctx_.track:
            ctx_.state.fwd_calls += 1
        return o

    @staticmethod
    def backward(ctx_, do: Tensor):
        q, kv = ctx_.saved_tensors
        group, state = ctx_.group, ctx_.state
        dkv = None
        if ctx_.track:
            off, c2 = group.offsets[ctx_.layer], group.widths[ctx_.layer]
            dkv = state.dkv_all[:, :, off:off + c2]
        dq, _ = ops.attn2_bwd(q, kv, do, ctx_.heads, ctx_.scale, dkv=dkv, pdl=ctx_.pdl)
        if ctx_.track:
            state.bwd_calls += 1
        return (dq if ctx_.needs_input_grad[0] else None), (do.new_zeros(1) if ctx_.has_token else None), None, None, None, None, None, None


class _PassState:
    """What one projection (one UNet pass) shares with its layers' backward: the fp32 dK|dV accumulator of all layers
    and the forward / backward call counts that guard the single dContext GEMM."""
    __slots__ = ("dkv_all", "fwd_calls", "bwd_calls")

    def __init__(self):
        self.dkv_all: Optional[Tensor] = None
        self.fwd_calls = 0
        self.bwd_calls = 0


class Attn2Group:
    """All ``attn2`` modules of one UNet, in forward-call order: owns the concatenated projection weights and, per
    forward pass, the projected context of every layer."""

    def __init__(self, attns):
        self.attns = list(attns)
        self.dims = [(a.heads, a.to_k.weight.shape[0] // a.heads, _padded_head_dim(a.to_k.weight.shape[0] // a.heads)) if hasattr(a, "heads")
                     else (1, a.to_k.weight.shape[0], a.to_k.weight.shape[0]) for a in self.attns]      # (H, d, padded d) per layer
        self.widths = [2 * h * dp for (h, d, dp) in self.dims]
        self.offsets = [0]
        for w in self.widths[:-1]:
            self.offsets.append(self.offsets[-1] + w)
        self.total = sum(self.widths)
        self.index = {id(a): i for i, a in enumerate(self.attns)}
        self._wcat: Optional[Tensor] = None
        self._wkey = None
        # WEAK reference to the context tensor object the cache belongs to: identity is the cache key (a dead object can
        # never alias a new one), and the previous pass's autograd graph -- which hangs off the student's context -- is not
        # kept alive across passes.  (A strong reference made the next pass re-use that graph's AccumulateGrad nodes, whose
        # streams belong to an earlier CUDA-graph capture: a second capture of the step then failed with
        # cudaErrorStreamCaptureIsolation.)
        self._ctx_ref = None
        self._ctx_key = None
        self._kv_all: Optional[Tensor] = None
        self._token: Optional[Tensor] = None
        self._supported: Optional[bool] = None
        self.state: Optional[_PassState] = None
        self.dctx_planes = 2                       # operand planes of dK|dV in the dContext GEMM (1 = throughput mode)
        self.sms = (torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
                    if torch.cuda.is_available() else 148)

    def weights(self, dtype, device) -> Tensor:
        key = (dtype, device, tuple(a.to_k.weight._version for a in self.attns), tuple(a.to_v.weight._version for a in self.attns),
               self.attns[0].to_k.weight.data_ptr())
        if self._wcat is None or self._wkey != key:
            parts = []
            for a, (h, d, dp) in zip(self.attns, self.dims):
                parts += [_pad_rows(a.to_k.weight.detach(), h, d, dp), _pad_rows(a.to_v.weight.detach(), h, d, dp)]
            self._wcat = torch.cat(parts, dim=0).to(device=device, dtype=dtype).contiguous()
            self._wkey = key
        return self._wcat

    def supported(self, dtype) -> bool:
        """LoRA layers on any to_k / to_v send the whole group back to the per-layer path (checked once per pass: the
        first layer's call; the answer is cached until the next projection)."""
        if self._supported is None:
            dc = self.attns[0].to_k.weight.shape[1]
            trainable = torch.is_grad_enabled() and any(a.to_k.weight.requires_grad or a.to_v.weight.requires_grad for a in self.attns)
            self._supported = dc % 8 == 0 and not trainable and all(
                a.to_k.weight.shape[1] == dc and getattr(a.to_k, "lora_layer", None) is None and
                getattr(a.to_v, "lora_layer", None) is None for a in self.attns)
        return self._supported

    def project(self, context: Tensor, dtype, layer: int = -1):
        """-> (kv_all [B*L, total], token or None) for the context tensor OBJECT the UNet passed down.  Recomputed when
        a different tensor (or an in-place modified one, a different dtype or grad mode) arrives; otherwise the
        pass's -- or, for a denoise loop that keeps feeding the same tensor, the loop's -- projection is reused.
        Under autograd a call from the FIRST attn2 layer always starts a new pass (own projection node, own dK|dV
        accumulator), even for the same tensor object: two passes must never share one gradient accumulator."""
        need_grad = torch.is_grad_enabled() and context.requires_grad
        self.weights(dtype, context.device)          # refreshes _wkey: a weight update between passes invalidates cached K|V
        key = (context._version, dtype, need_grad, self._wkey)
        new_pass = need_grad and layer == 0
        if not new_pass and self._ctx_ref is not None and self._ctx_ref() is context and self._ctx_key == key and self._kv_all is not None:
            tok, self._token = self._token, None     # only the first layer of a pass carries the token
            return self._kv_all, tok
        ctx = context.to(dtype)
        if not ctx.is_contiguous():
            ctx = ctx.contiguous()
        self._supported = None       # re-validated by the next pass
        kv_all, token = _ProjectAllFn.apply(ctx, self)
        self._ctx_ref, self._ctx_key = weakref.ref(context), key
        self._kv_all = kv_all
        self._token = None
        return kv_all, (token if need_grad else None)

    def release(self) -> None:
        self._ctx_ref = self._kv_all = self._token = self.state = None


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

    _warned = False

    def __init__(self, fallback=None, cache_kv: bool = False, group: Optional[Attn2Group] = None, own_projections: bool = False):
        self.fallback = fallback
        self.cache_kv = cache_kv
        self.group = group
        # to_q / to_out[0] on the library's GEMM (bias in the epilogue) as a PDL chain around the attention kernel; False
        # keeps the module's own nn.Linear calls (cuBLAS)
        self.own_projections = own_projections
        self._bias32: Optional[Tensor] = None
        self._bias_key = None
        self._pad_w = None          # (key, Wq padded [H*dp, C], Wo padded [C, H*dp]) for head sizes below 64

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

    def _linear(self, lin, x: Tensor, scale: float, is_out: bool) -> Tensor:
        """``lin(x)`` -- through the library's GEMM when ``lin`` is a plain frozen 16-bit Linear (no LoRA layer, weight in the
        activation dtype, TMA-compatible sizes), else the module's own forward (diffusers 0.23 LoRACompatibleLinear takes
        ``scale`` as a second positional argument only when a lora_layer is attached)."""
        w = lin.weight
        own = (self.own_projections and getattr(lin, "lora_layer", None) is None and w.dtype == x.dtype and
               not (torch.is_grad_enabled() and (w.requires_grad or (lin.bias is not None and lin.bias.requires_grad))) and
               w.shape[0] % 8 == 0 and w.shape[1] % 8 == 0 and w.is_contiguous() and w.data_ptr() % 16 == 0)
        if not own:
            return lin(x)
        b32 = None
        if lin.bias is not None:
            key = (lin.bias.data_ptr(), lin.bias._version)
            if self._bias32 is None or self._bias_key != key:
                self._bias32, self._bias_key = lin.bias.detach().float().contiguous(), key
            b32 = self._bias32
        return _FrozenLinearFn.apply(x, w.detach(), b32)

    def _padded_projections(self, attn, dt, d: int, dp: int):
        """Zero-padded copies of the frozen to_q / to_out[0] weights (built once per weight version), or None when those
        modules are not plain frozen Linears -- then the layer keeps its true head size and the warp-MMA kernels."""
        lq, lo = attn.to_q, attn.to_out[0]
        for lin in (lq, lo):
            if getattr(lin, "lora_layer", None) is not None or lin.weight.dtype != dt or \
                    (torch.is_grad_enabled() and (lin.weight.requires_grad or (lin.bias is not None and lin.bias.requires_grad))):
                return None
        key = (lq.weight.data_ptr(), lq.weight._version, lo.weight.data_ptr(), lo.weight._version, dt)
        if self._pad_w is None or self._pad_w[0] != key:
            self._pad_w = (key, _pad_rows(lq.weight.detach(), attn.heads, d, dp).contiguous(),
                           _pad_cols(lo.weight.detach(), attn.heads, d, dp).contiguous())
        return self._pad_w[1], self._pad_w[2]

    def _supported(self, attn, hidden_states, encoder_hidden_states, attention_mask) -> bool:
        if encoder_hidden_states is None or attention_mask is not None:
            return False
        if hidden_states.dim() != 3 or hidden_states.dtype not in (torch.bfloat16, torch.float16):
            return False
        if getattr(attn.to_k, "lora_layer", None) is not None or getattr(attn.to_v, "lora_layer", None) is not None:
            return False
        if torch.is_grad_enabled() and (attn.to_k.weight.requires_grad or attn.to_v.weight.requires_grad):
            return False   # the fused path never forms dWk / dWv (frozen-UNet contract): trainable projections -> stock processor
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
        if encoder_hidden_states is not None:
            ops.require_cuda(hidden_states, encoder_hidden_states)     # cross-attention on host tensors: no CPU path, ever
        if not self._supported(attn, hidden_states, encoder_hidden_states, attention_mask):
            if encoder_hidden_states is not None and hidden_states.is_cuda and not PEAAttn2Processor._warned:
                PEAAttn2Processor._warned = True   # never silent: a cross-attention call is leaving the fused kernels
                warnings.warn("PEAAttn2Processor: cross-attention call outside the fused kernel's contract (mask / LoRA on "
                              f"to_k,to_v / > {MAX_KEYS} keys / head size / dtype {hidden_states.dtype}); using the stock processor")
            if self.fallback is None:
                raise ops._lib.PeaError("PEAAttn2Processor: call outside the fused kernel's contract and no fallback processor set")
            return self.fallback(attn, hidden_states, encoder_hidden_states=encoder_hidden_states,
                                 attention_mask=attention_mask, temb=temb, scale=scale)
        dt = hidden_states.dtype
        d = attn.to_q.weight.shape[0] // attn.heads
        dp = _padded_head_dim(d)
        pad = self._padded_projections(attn, dt, d, dp) if dp != d else None
        grouped = self.group is not None and id(attn) in self.group.index and self.group.supported(dt)
        if dp != d and (pad is None or not grouped):
            dp, pad = d, None          # per-layer path / non-plain projections: true head size (warp-MMA kernels)
            grouped = grouped and self.group.dims[self.group.index[id(attn)]][2] == d
        if pad is not None:
            q = torch.nn.functional.linear(hidden_states, pad[0])      # [B, N, H * 64]: the padded dims are exact zeros
        else:
            q = self._linear(attn.to_q, hidden_states, scale, False)
        if grouped:
            # all layers' K|V from ONE projection GEMM per forward pass (and per denoise loop when the context is reused)
            B, L, _ = encoder_hidden_states.shape
            i = self.group.index[id(attn)]
            kv_all, token = self.group.project(encoder_hidden_states, dt, i)
            off, c2 = self.group.offsets[i], self.group.widths[i]
            kv = kv_all.view(B, L, self.group.total)[:, :, off:off + c2]
            o = _Attn2GroupedFn.apply(q, token, kv, self.group, i, attn.heads, float(attn.scale), self.own_projections and pad is None)
            if pad is not None:
                o = torch.nn.functional.linear(o, pad[1], attn.to_out[0].bias)
            else:
                o = self._linear(attn.to_out[0], o, scale, True)
            return attn.to_out[1](o)
        context = encoder_hidden_states.to(dt)
        if not context.is_contiguous():
            context = context.contiguous()
        wkv = self._weights(attn, dt)
        kv = None
        if self.cache_kv and not (torch.is_grad_enabled() and context.requires_grad):
            # identity of the context TENSOR OBJECT (held strongly, so its address cannot be recycled) + its version
            key = (id(encoder_hidden_states), encoder_hidden_states._version, tuple(context.shape), dt)
            self._kv_ctx = encoder_hidden_states
            if self._kv is None or self._kv_key != key:
                B, L, Dc = context.shape
                self._kv = torch.empty(B, L, wkv.shape[0], device=q.device, dtype=dt)
                ops.gemm([context.reshape(B * L, Dc)], [wkv], B * L, wkv.shape[0], Dc, out0=self._kv.view(B * L, -1))
                self._kv_key = key
            kv = self._kv
        o = _Attn2Fn.apply(q, context, wkv, attn.heads, float(attn.scale), kv)
        o = self._linear(attn.to_out[0], o, scale, True)
        return attn.to_out[1](o)


def install(unet, freeze: bool = True, cache_kv: bool = False, group_projection: bool = True,
            own_projections: bool = False) -> Dict[str, object]:
    """``unet.set_attn_processor({...})`` with ``PEAAttn2Processor`` on every ``*.attn2`` module and the existing
    processor kept everywhere else (diffusers requires the dict to cover every attention module).

    ``group_projection=True`` (default) projects the context for ALL attn2 layers with one GEMM per forward pass (and
    forms dContext with one split-K GEMM per backward pass); with ``cache_kv`` the projection also survives across
    passes while the same context tensor is fed (denoise loops).  ``freeze=True`` also marks the UNet's parameters ``requires_grad_(False)``: the reference hands only the adapter's
    parameters to the optimizer (train_sdxl_zh.py:166-168) yet lets autograd form and discard every UNet weight
    gradient; skipping them does not change any result.
    """
    current = dict(unet.attn_processors)
    group = None
    if group_projection:
        # attn2 modules in registration (= forward-call) order; one projection GEMM per pass serves all of them
        mods = dict(unet.named_modules())
        attns = [mods[name[:-len(".processor")]] for name in current if name.endswith("attn2.processor")]
        if attns and len({a.to_k.weight.shape[1] for a in attns}) == 1:
            group = Attn2Group(attns)
    procs = {}
    for name, proc in current.items():
        procs[name] = (PEAAttn2Processor(fallback=proc, cache_kv=cache_kv, group=group, own_projections=own_projections)
                       if name.endswith("attn2.processor") else proc)
    unet.set_attn_processor(procs)
    if freeze:
        unet.requires_grad_(False)
    return procs
