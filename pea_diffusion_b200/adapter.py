"""The PEA adapter -- drop-in ``nn.Module``s with the reference's constructors, forward signatures and checkpoint keys.

Reference classes (each script re-declares its own ``MLP``):
    MLP            train_sdxl_zh.py:43-67 (== tests/test_sdxl_zh.py:59-84, ``MLP_plus`` tests/test_sdxl_zh_lcm.py:89-114)
    MLP_SD         train_sd_zh.py:41-56 (3 Linears) and tests/test_sd_zh.py:25-44 (5 Linears, hidden 3072)
    MLP_Legacy     tests/test_sdxl_zh_lcm.py:64-87 (fc1/fc2/fc3 with bias)

Parameters live in ordinary ``nn.LayerNorm`` / ``nn.Linear`` submodules named exactly as in the reference, so
``state_dict()`` / ``load_state_dict()`` / ``.half()`` behave identically; only ``forward`` differs: it runs the whole
LayerNorm -> Linear/GELU chain -> {mean-pool, GELU -> fc} graph, forward and backward, on the tcgen05 GEMM of
``libpea_b200`` with bias / GELU / GELU' / pooling-gradient fused into the GEMM epilogues.

Precision modes (``precision=``):
    "fp32"  (default for fp32 parameters)  operands are carried as TWO bf16 planes (hi + residual) and the three most
            significant cross products are accumulated in fp32 in TMEM: ~2^-16 relative, inside the 1e-4 parity
            tolerance of the north star, at 3x the MMA count of plain bf16 (the adapter is launch-latency bound).
    "bf16" / "fp16"  single-plane operands, fp32 accumulation (the reference trains in fp16 autocast-like mode).
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import ops

Tensor = torch.Tensor

# Every launch of the adapter's forward / backward chain carries PEA_LAUNCH_PDL (include/pea_b200.h): the chain is ~30
# launches of a few microseconds each, so the launch-to-launch gap is a third of its time.  PEA_ADAPTER_PDL=0 disables (A/B).
PDL_CHAIN = os.environ.get("PEA_ADAPTER_PDL", "1") != "0"


def _mode(precision: Optional[str], param_dtype: torch.dtype) -> Tuple[int, torch.dtype]:
    if precision is None:
        precision = {torch.float32: "fp32", torch.bfloat16: "bf16", torch.float16: "fp16"}[param_dtype]
    if precision == "fp32":
        return 2, torch.bfloat16
    if precision == "fp32x3":
        return 3, torch.bfloat16
    if precision == "bf16":
        return 1, torch.bfloat16
    if precision == "fp16":
        return 1, torch.float16
    raise ValueError(f"unknown precision {precision!r}")


class _AdapterFn(torch.autograd.Function):
    """LN -> [Linear (+bias) -> GELU]* -> Linear (+bias) = x ;  optional head x2 = Linear(GELU(x)) + b ;
    optional pooled x1 = mean_tokens(x (+ input)).

    Flat tensor arguments: x, ln_w, ln_b, then (W_i, b_i|None) per chain layer, then (W_head, b_head) if ``head``.
    """

    @staticmethod
    def forward(ctx, nsplit: int, plane_dt: torch.dtype, n_chain: int, head: bool, pooled: bool, residual: bool,
                x: Tensor, ln_w: Tensor, ln_b: Tensor, sink, *wb: Optional[Tensor]):
        ops.require_cuda(x, ln_w, ln_b)
        B, L, Din = x.shape
        M = B * L
        dev = x.device
        f32 = torch.float32
        x2d = x.detach().reshape(M, Din)
        if not x2d.is_contiguous():
            x2d = x2d.contiguous()
        Ws = [wb[2 * i] for i in range(n_chain + int(head))]
        bs = [wb[2 * i + 1] for i in range(n_chain + int(head))]
        w_planes = ops.split_planes_multi([w.detach() for w in Ws], nsplit, plane_dt, pdl=PDL_CHAIN)   # one launch
        b_f32 = [None if b is None else b.detach().to(f32).contiguous() for b in bs]

        h, mean, rstd = ops.layernorm_fwd(x2d, ln_w.detach().to(f32).contiguous(), ln_b.detach().to(f32).contiguous(), 1e-5,
                                          nsplit, plane_dt, pdl=PDL_CHAIN)
        inputs: List[List[Tensor]] = []   # A operand (planes) of every GEMM, kept for dW
        preacts: List[Tensor] = []        # fp32 pre-activation of every chain layer, kept for GELU'
        for i in range(n_chain):
            inputs.append(h)
            N = Ws[i].shape[0]
            a = torch.empty(M, N, device=dev, dtype=f32)
            last = i == n_chain - 1
            nxt = None
            if not last or head:
                nxt = [torch.empty(M, N, device=dev, dtype=plane_dt) for _ in range(nsplit)]
            ops.gemm(h, w_planes[i], M, N, Ws[i].shape[1], bias=b_f32[i], out0=a, out1=nxt, act=True, pdl=PDL_CHAIN)
            preacts.append(a)
            h = nxt
        xlast = preacts[-1]
        outs = []
        if pooled:
            x1 = ops.mean_tokens(xlast, B, L, res=x2d if residual else None, pdl=PDL_CHAIN)
            outs.append(x1.to(x.dtype))
        if head:
            inputs.append(h)
            N = Ws[n_chain].shape[0]
            x2 = torch.empty(M, N, device=dev, dtype=x.dtype)
            ops.gemm(h, w_planes[n_chain], M, N, Ws[n_chain].shape[1], bias=b_f32[n_chain], out0=x2, pdl=PDL_CHAIN)
            outs.append(x2.view(B, L, N))
        else:
            outs.append(xlast.to(x.dtype).view(B, L, -1))
        ctx.cfg = (nsplit, plane_dt, n_chain, head, pooled, residual, B, L)
        ctx.saved = (x2d, mean, rstd, inputs, preacts, w_planes, ln_w, ln_b, Ws, bs)
        ctx.sink = sink
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts: Optional[Tensor]):
        nsplit, plane_dt, n_chain, head, pooled, residual, B, L = ctx.cfg
        x2d, mean, rstd, inputs, preacts, w_planes, ln_w, ln_b, Ws, bs = ctx.saved
        sink = ctx.sink     # optional gradient sink (train.GradAllReducer): dW / db are written straight into its flat
        #                     communication buffer and the bucket's all-reduce starts while the rest of this backward runs
        M = B * L
        dev = x2d.device
        f32 = torch.float32
        gouts = list(gouts)
        g_pool = gouts.pop(0) if pooled else None
        g_main = gouts.pop(0)
        n_lin = n_chain + int(head)
        dWs: List[Optional[Tensor]] = [None] * n_lin
        dbs: List[Optional[Tensor]] = [None] * n_lin
        rowadd = None if g_pool is None else g_pool.detach().to(f32).contiguous()

        # fp32 accumulators of the reduction kernels (bias gradients, dgamma, dbeta).  With a sink whose slot is fp32 the
        # kernels accumulate straight into the (zeroed) slot; everything else is carved from ONE zero-filled buffer.
        Dln = x2d.shape[1]
        need = 2 * Dln + sum(Ws[i].shape[0] for i in range(n_lin) if bs[i] is not None and bs[i].requires_grad)
        zpool, zoff = torch.zeros(need, device=dev, dtype=f32), [0]

        def accumulator(prm: Tensor, n: int) -> Tuple[Tensor, Optional[Tensor]]:
            """-> (fp32 accumulator of n zeros, the sink slot of `prm` or None)"""
            slot = sink.slot(prm) if sink is not None else None
            if slot is not None and slot.dtype == f32 and slot.is_contiguous():
                return slot.view(-1).zero_(), slot
            z = zpool[zoff[0]:zoff[0] + n]
            zoff[0] += n
            return z, slot

        def deliver(prm: Tensor, acc: Tensor, slot: Optional[Tensor]) -> Optional[Tensor]:
            """-> the gradient to hand to autograd (None when it went to the sink)"""
            if slot is None:
                return acc.to(prm.dtype).view_as(prm)
            if slot.data_ptr() != acc.data_ptr():
                slot.copy_(acc.view_as(slot))
            sink.mark_ready(prm)
            return None

        def weight_grads(i: int, g_planes: Sequence[Tensor], g_full: Optional[Tensor]):
            W = Ws[i]
            if W.requires_grad:
                slot = sink.slot(W) if sink is not None else None
                dW = slot if slot is not None else torch.empty(W.shape, device=dev, dtype=W.dtype)
                # dW[o, k] = sum_r g[r, o] h[r, k]: both operands are stored [rows, *] -> MN-major
                ops.gemm(g_planes, inputs[i], W.shape[0], W.shape[1], M, a_kmajor=False, b_kmajor=False, out0=dW,
                         pdl=PDL_CHAIN)
                if slot is None:
                    dWs[i] = dW
                else:
                    sink.mark_ready(W)
            if bs[i] is not None and bs[i].requires_grad:
                db, slot = accumulator(bs[i], W.shape[0])
                ops.colsum_accum(g_full if g_full is not None else g_planes[0], db, pdl=PDL_CHAIN)
                if g_full is None:
                    for extra in g_planes[1:]:
                        ops.colsum_accum(extra, db, pdl=PDL_CHAIN)
                dbs[i] = deliver(bs[i], db, slot)

        # ---- gradient w.r.t. the last chain pre-activation x -------------------------------------------------
        if head:
            if g_main is None:
                g_main = torch.zeros(M, Ws[n_chain].shape[0], device=dev, dtype=x2d.dtype)
            g2d = g_main.detach().reshape(M, -1)
            g_planes = ops.split_planes(g2d, nsplit, plane_dt, pdl=PDL_CHAIN)
            weight_grads(n_chain, g_planes, g2d)
            N = Ws[n_chain].shape[1]
            nxt = [torch.empty(M, N, device=dev, dtype=plane_dt) for _ in range(nsplit)]
            full = torch.empty(M, N, device=dev, dtype=f32) if bs[n_chain - 1] is not None else None
            # dx = (g2 @ Wfc) * GELU'(x) + dx1[b] / L
            ops.gemm(g_planes, w_planes[n_chain], M, N, Ws[n_chain].shape[0], b_kmajor=False, aux=preacts[-1], rowadd=rowadd,
                     rowadd_group=L, rowadd_scale=1.0 / L, out0=full, out1=nxt, pdl=PDL_CHAIN)
            g_planes, g_full = nxt, full
        else:
            g = g_main.detach().reshape(M, -1).to(f32)
            if rowadd is not None:
                g = g + (rowadd / L).repeat_interleave(L, dim=0)
            g_planes, g_full = ops.split_planes(g.contiguous(), nsplit, plane_dt, pdl=PDL_CHAIN), g

        # ---- chain, last to first -------------------------------------------------------------------------------
        dh0 = None
        for i in range(n_chain - 1, -1, -1):
            weight_grads(i, g_planes, g_full)
            N = Ws[i].shape[1]
            if i > 0:
                nxt = [torch.empty(M, N, device=dev, dtype=plane_dt) for _ in range(nsplit)]
                full = torch.empty(M, N, device=dev, dtype=f32) if bs[i - 1] is not None else None
                ops.gemm(g_planes, w_planes[i], M, N, Ws[i].shape[0], b_kmajor=False, aux=preacts[i - 1], out0=full, out1=nxt,
                         pdl=PDL_CHAIN)
                g_planes, g_full = nxt, full
            else:
                dh0 = torch.empty(M, N, device=dev, dtype=f32)
                ops.gemm(g_planes, w_planes[0], M, N, Ws[0].shape[0], b_kmajor=False, out0=dh0, pdl=PDL_CHAIN)

        # ---- LayerNorm ------------------------------------------------------------------------------------------
        want_dx = ctx.needs_input_grad[6]
        (dgamma, slot_g), (dbeta, slot_b) = accumulator(ln_w, Dln), accumulator(ln_b, Dln)
        dx = ops.layernorm_bwd(x2d, dh0, ln_w.detach().to(f32).contiguous(), mean, rstd, dgamma, dbeta, want_dx, pdl=PDL_CHAIN)
        if want_dx:
            if residual and rowadd is not None:
                dx = dx + (rowadd / L).repeat_interleave(L, dim=0)
            dx = dx.to(x2d.dtype).view(B, L, -1)
        flat: List[Optional[Tensor]] = []
        for i in range(n_lin):
            flat += [dWs[i], dbs[i]]
        g_lnw, g_lnb = deliver(ln_w, dgamma, slot_g), deliver(ln_b, dbeta, slot_b)
        return (None, None, None, None, None, None, dx if want_dx else None, g_lnw, g_lnb, None, *flat)


def _run(module: nn.Module, x: Tensor, chain: Sequence[nn.Linear], head: Optional[nn.Linear], pooled: bool, residual: bool):
    if x.dim() != 3:
        raise ValueError(f"adapter input must be [B, L, in_dim], got {tuple(x.shape)}")
    p_dtype = chain[0].weight.dtype
    nsplit, plane_dt = _mode(getattr(module, "precision", None), p_dtype)
    if x.dtype != p_dtype:
        raise RuntimeError(f"mat1 and mat2 must have the same dtype, but got {x.dtype} and {p_dtype}")  # nn.Linear's own error
    wb: List[Optional[Tensor]] = []
    for lin in list(chain) + ([head] if head is not None else []):
        wb += [lin.weight, lin.bias]
    return _AdapterFn.apply(nsplit, plane_dt, len(chain), head is not None, pooled, residual, x, module.layernorm.weight,
                            module.layernorm.bias, getattr(module, "_grad_sink", None), *wb)


class MLP(nn.Module):
    """SDXL adapter: ``MLP(in_dim, out_dim, hidden_dim, out_dim1, use_residual)``; ``forward(x[B,L,in]) -> (x1[B,out],
    x2[B,L,out1])`` (train_sdxl_zh.py:43-67).  Keys: layernorm.{weight,bias}, projector.{0,2,4}.weight, fc.{weight,bias}."""

    def __init__(self, in_dim=1024, out_dim=1280, hidden_dim=2048, out_dim1=2048, use_residual=True, precision: Optional[str] = None):
        super().__init__()
        if use_residual:
            assert in_dim == out_dim
        self.layernorm = nn.LayerNorm(in_dim)
        self.projector = nn.Sequential(
            nn.Linear(in_dim, hidden_dim, bias=False), nn.GELU(),
            nn.Linear(hidden_dim, hidden_dim, bias=False), nn.GELU(),
            nn.Linear(hidden_dim, out_dim, bias=False),
        )
        self.fc = nn.Linear(out_dim, out_dim1)
        self.use_residual = use_residual
        self.precision = precision

    def forward(self, x):
        x1, x2 = _run(self, x, [m for m in self.projector if isinstance(m, nn.Linear)], self.fc, True, self.use_residual)
        return x1, x2


MLP_plus = MLP  # tests/test_sdxl_zh_lcm.py:89-114 names the same module MLP_plus


class MLP_SD(nn.Module):
    """SD1.5 adapter: ``MLP(in_dim, out_dim, hidden_dim)`` -> ``forward(x) -> x[B,L,out]`` (train_sd_zh.py:41-56);
    ``num_linear=5`` gives the inference variant of tests/test_sd_zh.py:25-44 (keys projector.{0,2,4,6,8}.weight)."""

    def __init__(self, in_dim=1024, out_dim=768, hidden_dim=2048, num_linear: int = 3, precision: Optional[str] = None):
        super().__init__()
        self.layernorm = nn.LayerNorm(in_dim)
        dims = [in_dim] + [hidden_dim] * (num_linear - 1) + [out_dim]
        mods: List[nn.Module] = []
        for i in range(num_linear):
            mods.append(nn.Linear(dims[i], dims[i + 1], bias=False))
            if i + 1 < num_linear:
                mods.append(nn.GELU())
        self.projector = nn.Sequential(*mods)
        self.precision = precision

    def forward(self, x):
        return _run(self, x, [m for m in self.projector if isinstance(m, nn.Linear)], None, False, False)[0]


class MLP_Legacy(nn.Module):
    """Legacy biased adapter (tests/test_sdxl_zh_lcm.py:64-87): LN -> fc1 -> GELU -> fc2 = x; x2 = fc3(GELU(x));
    x1 = mean(x).  Keys: layernorm.*, fc1.*, fc2.*, fc3.*."""

    def __init__(self, in_dim=1024, out_dim=1280, hidden_dim=2048, out_dim1=2048, use_residual=True, precision: Optional[str] = None):
        super().__init__()
        if use_residual:
            assert in_dim == out_dim
        self.layernorm = nn.LayerNorm(in_dim)
        self.fc1 = nn.Linear(in_dim, hidden_dim)
        self.fc2 = nn.Linear(hidden_dim, out_dim)
        self.fc3 = nn.Linear(out_dim, out_dim1)
        self.use_residual = use_residual
        self.precision = precision

    def forward(self, x):
        x1, x2 = _run(self, x, [self.fc1, self.fc2], self.fc3, True, self.use_residual)
        return x1, x2
