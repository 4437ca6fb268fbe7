"""Teacher/student KD loss of PEA-Diffusion as ONE fused CUDA pass (loss and d(student) together).

Reference arithmetic:
    train_sdxl_zh.py:400-405   L0 = mean_b(mean_chw((eps_S - eps)^2   * zh_b))
    train_sdxl_zh.py:416-421   L1 = mean_b(mean_chw((eps_S - eps_T)^2 * (1 - zh_b)))
    train_sdxl_zh.py:423-441   L2 = sum_taps mean_b(mean_chw((f_T - f_S)^2 * (1 - zh_b)))      L = L0 + L1 + 0.1 * L2
    train_sd_zh.py:246-268     a feature tap holding NaN/Inf is skipped (``skip_nonfinite``)

The reference makes ~5 HBM passes forward and ~4 backward per tap and three ``.item()`` host syncs; here the forward
call reads every student/teacher tensor once, writes d(student) once and leaves the four loss scalars on the device.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

from . import ops

Tensor = torch.Tensor


class _KDLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feature_weight: float, skip_nonfinite: bool, noise: Tensor, pred_t: Tensor, zh: Tensor, nfeat: int,
                pred: Tensor, grad_prescale: float, *feats: Tensor):
        fs, ft = feats[:nfeat], feats[nfeat:]
        dt = pred.dtype
        B = pred.shape[0]
        dev = pred.device
        w_zh = zh.to(device=dev, dtype=torch.float32).contiguous()
        w_en = (1.0 - w_zh).contiguous()
        need_grad = ctx.needs_input_grad[6] or any(ctx.needs_input_grad[8:8 + nfeat])

        def same(t):  # targets follow the student's dtype (the reference upcasts both sides to fp32)
            return t.detach().to(dtype=dt).contiguous()

        taps = ops.KdTaps()
        grads: List[Tensor] = []
        pred_c = pred.detach().contiguous()
        g = torch.empty_like(pred_c) if need_grad else None
        grads.append(g)
        taps.add(pred_c, same(noise), w_zh, 0, grad=g, teacher2=same(pred_t), weight2=w_en, group2=1)
        for s, t in zip(fs, ft):
            if s.shape != t.shape or s.shape[0] != B:
                raise ValueError(f"KD tap shapes differ: student {tuple(s.shape)} teacher {tuple(t.shape)}")
            s_c = s.detach().to(dtype=dt).contiguous()
            g = torch.empty_like(s_c) if need_grad else None
            grads.append(g)
            taps.add(s_c, same(t), w_en, 2, grad=g)
        # d(student) is produced in THIS pass, already multiplied by `grad_prescale` -- the loss scale the caller will
        # back-propagate -- before it is rounded to the student dtype.  In fp16 the unscaled feature-KD coefficient
        # 0.1 * 2 / (numel * B) ~ 1e-8 is below the smallest fp16 subnormal; scaling after the rounding (in backward)
        # could not recover it.  The reference gets the same effect from DeepSpeed's loss scaler (train_sdxl_zh.sh:89).
        out = ops.kd_mse_fwd_bwd(taps, B, dt, feature_weight, float(grad_prescale), skip_nonfinite, dev)
        ctx.taps, ctx.B, ctx.dt, ctx.dev, ctx.out, ctx.grads, ctx.nfeat = taps, B, dt, dev, out, grads, nfeat
        ctx.inv_prescale = 1.0 / float(grad_prescale)
        ctx.in_dtypes = [pred.dtype] + [s.dtype for s in fs]
        ctx.mark_non_differentiable(out)
        return out[3].clone(), out

    @staticmethod
    def backward(ctx, g_total, _g_out):
        # d(student) was produced for upstream == grad_prescale; rescale by upstream / grad_prescale on the device (a
        # no-op kernel exit when that is 1 and no tap was skipped) -- no host sync.
        up = g_total.detach().to(torch.float32).reshape(1).contiguous()
        ops.kd_scale_grads(ctx.taps, ctx.B, ctx.dt, up, ctx.out, ctx.dev, inv_prescale=ctx.inv_prescale)
        gs = [g if g is None or g.dtype == d else g.to(d) for g, d in zip(ctx.grads, ctx.in_dtypes)]
        return (None, None, None, None, None, None, gs[0], None, *gs[1:], *([None] * ctx.nfeat))


def kd_loss(noise_pred: Tensor, noise: Tensor, noise_pred_teacher: Tensor, feats_student: Sequence[Tensor],
            feats_teacher: Sequence[Tensor], zh_or_not: Tensor, feature_weight: float = 0.1,
            skip_nonfinite: bool = False, grad_prescale: float = 1.0) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """-> (total, L0, L1, L2) as 0-dim fp32 device tensors; ``total.backward()`` yields d(noise_pred), d(feats_student).

    ``grad_prescale``: the loss scale the caller is going to back-propagate (``(total * S).backward()``); pass it when
    training in fp16 so the fused pass scales the gradient before rounding it.  The returned losses are unscaled, and
    any other upstream gradient still gives the exact result (one extra rescale pass).

    ``zh_or_not`` is the batch's int64 [B] language flag (utils/custom_dataset_sdxl.py collate, train_sdxl_zh.py:399).
    Tensors are NCHW [B, ...]; teacher-side tensors get no gradient (``no_grad`` in the reference, :410-415).
    """
    feats_student, feats_teacher = list(feats_student), list(feats_teacher)
    if len(feats_student) != len(feats_teacher):
        raise ValueError("feats_student / feats_teacher length mismatch")
    total, out = _KDLossFn.apply(float(feature_weight), bool(skip_nonfinite), noise, noise_pred_teacher, zh_or_not,
                                 len(feats_student), noise_pred, float(grad_prescale), *feats_student, *feats_teacher)
    return total, out[0], out[1], out[2]
