"""ORACLE -- test infrastructure, NOT product code.

CPU restatement (plain PyTorch, fp32/fp64, autograd for gradients) of the
PEA-Diffusion KD hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
module; nothing under ``pea_diffusion_b200/`` does.

PARITY STATUS: *partially pinned*.
  * Adapter (rows A1/A2/A4 of SURVEY.md section 8a): pinned.  The reference's
    ``MLP`` classes are pure torch and were executed verbatim in the build
    container (AST-extracted from the cited lines by
    ``tests/golden/make_golden.py``); their outputs and gradients are committed
    under ``tests/golden/`` and this restatement is checked against them.
  * KD loss (A8/A9): pinned the same way -- ``make_golden.py`` executes the
    reference's own loss statements (train_sdxl_zh.py:400-441,
    train_sd_zh.py:218-272) extracted by line range.
  * attn2 (A5/A6): PARITY UNPINNED by the reference.  The arithmetic lives in
    diffusers==0.23.0 ``AttnProcessor2_0`` + torch==2.0.1 SDPA
    (requirements.txt:14,24), neither vendored nor installable offline, and
    the reference holds no golden vectors.  The restatement below follows the
    published algorithm and is cross-checked against
    torch's own ``F.scaled_dot_product_attention`` and an fp64 explicit
    softmax(QK^T)V.

Every function cites the reference lines it follows.
"""
from __future__ import annotations

import math
from typing import Dict, List, Mapping, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# =============================================================================
# A1 / A2 / A4 -- the PEA adapter
# =============================================================================
def _projector_weights(sd: Mapping[str, Tensor]) -> List[Tensor]:
    """``projector`` is an nn.Sequential with bias-free Linears at even indices and exact GELUs between
    (train_sdxl_zh.py:49-55, train_sd_zh.py:45-51, tests/test_sd_zh.py:29-39)."""
    idx = sorted(int(k.split(".")[1]) for k in sd if k.startswith("projector.") and k.endswith(".weight"))
    return [sd[f"projector.{i}.weight"] for i in idx]


def _projector(sd: Mapping[str, Tensor], h: Tensor) -> Tensor:
    ws = _projector_weights(sd)
    for i, w in enumerate(ws):
        h = F.linear(h, w)
        if i + 1 < len(ws):
            h = F.gelu(h)  # nn.GELU() == exact erf form
    return h


def adapter_sdxl(sd: Mapping[str, Tensor], x: Tensor, use_residual: bool = False) -> Tuple[Tensor, Tensor]:
    """SDXL adapter forward, train_sdxl_zh.py:58-67 (== MLP_plus, tests/test_sdxl_zh_lcm.py:104-113,
    == tests/test_sdxl_zh.py:59-84).

    LN(eps 1e-5, affine) -> projector -> x ; x2 = fc(GELU(x)) ; x1 = mean over ALL tokens (dim 1) of the
    pre-GELU x (plus the input when ``use_residual``).  Returns (x1 [B,out], x2 [B,L,out1]).
    """
    in_dim = x.shape[-1]
    h = F.layer_norm(x, (in_dim,), sd["layernorm.weight"], sd["layernorm.bias"], 1e-5)
    h = _projector(sd, h)
    x2 = F.linear(F.gelu(h), sd["fc.weight"], sd["fc.bias"])
    if use_residual:
        h = h + x
    return h.mean(dim=1), x2


def adapter_sd(sd: Mapping[str, Tensor], x: Tensor) -> Tensor:
    """SD1.5 adapter forward: train_sd_zh.py:53-56 (3 Linears) and tests/test_sd_zh.py:41-44 (5 Linears).
    Single output [B,L,out]."""
    h = F.layer_norm(x, (x.shape[-1],), sd["layernorm.weight"], sd["layernorm.bias"], 1e-5)
    return _projector(sd, h)


def adapter_legacy(sd: Mapping[str, Tensor], x: Tensor, use_residual: bool = False) -> Tuple[Tensor, Tensor]:
    """Legacy biased adapter, tests/test_sdxl_zh_lcm.py:75-87: LN -> fc1 -> GELU -> fc2 -> x ; x2 = fc3(GELU(x))."""
    h = F.layer_norm(x, (x.shape[-1],), sd["layernorm.weight"], sd["layernorm.bias"], 1e-5)
    h = F.gelu(F.linear(h, sd["fc1.weight"], sd["fc1.bias"]))
    h = F.linear(h, sd["fc2.weight"], sd["fc2.bias"])
    x2 = F.linear(F.gelu(h), sd["fc3.weight"], sd["fc3.bias"])
    if use_residual:
        h = h + x
    return h.mean(dim=1), x2


def adapter_init(in_dim: int, out_dim: int, hidden_dim: int, out_dim1: Optional[int] = None, n_linear: int = 3,
                 seed: int = 0, dtype=torch.float32) -> Dict[str, Tensor]:
    """Random-init state dict with the reference's key names and shapes: LayerNorm at
    (1, 0), Linear weights/biases ~ U(+-1/sqrt(fan_in)) as nn.Linear draws them, from a private seeded generator
    (so full-size weights can be regenerated from a seed instead of being stored in fixtures)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {"layernorm.weight": torch.ones(in_dim), "layernorm.bias": torch.zeros(in_dim)}
    dims = [in_dim] + [hidden_dim] * (n_linear - 1) + [out_dim]

    def kaiming(o, i):
        bound = 1.0 / math.sqrt(i)
        return (torch.rand(o, i, generator=g) * 2 - 1) * bound

    for j in range(n_linear):
        sd[f"projector.{2 * j}.weight"] = kaiming(dims[j + 1], dims[j])
    if out_dim1 is not None:
        sd["fc.weight"] = kaiming(out_dim1, out_dim)
        sd["fc.bias"] = (torch.rand(out_dim1, generator=g) * 2 - 1) / math.sqrt(out_dim)
    return {k: v.to(dtype) for k, v in sd.items()}


# =============================================================================
# A3 -- classifier-free-guidance dropout select
# =============================================================================
def cfg_select(prompt_mask: Tensor, ctx_uncond: Tensor, ctx_cond: Tensor) -> Tensor:
    """train_sdxl_zh.py:392-395 / train_sd_zh.py:210-213: ``where(mask[n,1,1], uncond, cond)``; mask = rand(B) < 0.1."""
    return torch.where(prompt_mask.view(-1, 1, 1), ctx_uncond, ctx_cond)


# =============================================================================
# N4 -- input side: offset noise + DDPM add_noise (train_sdxl_zh.py:311-323)
# =============================================================================
def add_noise(latents: Tensor, noise_raw: Tensor, offs: Tensor, noise_offset: float, alphas_cumprod: Tensor, timesteps: Tensor,
              unet_dtype: torch.dtype) -> Tuple[Tensor, Tensor]:
    """``noise = randn + noise_offset * randn(B, C, 1, 1)`` (:311-315), cast to the UNet dtype (:317), then diffusers'
    ``DDPMScheduler.add_noise``: ``sqrt(acp[t]) * latents + sqrt(1 - acp[t]) * noise`` (:322), cast (:323)."""
    noise = (noise_raw + noise_offset * offs.view(offs.shape[0], offs.shape[1], 1, 1)).to(unet_dtype)
    a = alphas_cumprod.to(torch.float32)[timesteps]
    sa, sb = a.sqrt().view(-1, 1, 1, 1), (1 - a).sqrt().view(-1, 1, 1, 1)
    noisy = (sa * latents.float() + sb * noise.float()).to(unet_dtype)
    return noise, noisy


# =============================================================================
# A5 / A6 -- attn2 (diffusers AttnProcessor2_0 restated; reference call sites train_sdxl_zh.py:397,415)
# =============================================================================
def attn2_core(q: Tensor, k: Tensor, v: Tensor, heads: int, explicit: bool = False) -> Tensor:
    """softmax(q k^T d^-1/2) v per head.  q [B,N,C], k/v [B,L,C] -> [B,N,C].  No mask, no dropout, non-causal."""
    b, n, c = q.shape
    d = c // heads
    qh = q.view(b, n, heads, d).transpose(1, 2)
    kh = k.view(b, -1, heads, d).transpose(1, 2)
    vh = v.view(b, -1, heads, d).transpose(1, 2)
    if explicit:
        s = torch.matmul(qh, kh.transpose(-1, -2)) * (d ** -0.5)
        o = torch.matmul(torch.softmax(s, dim=-1), vh)
    else:
        o = F.scaled_dot_product_attention(qh, kh, vh, attn_mask=None, dropout_p=0.0, is_causal=False)
    return o.transpose(1, 2).reshape(b, n, c)


def attn2(hidden: Tensor, ctx: Tensor, wq: Tensor, wk: Tensor, wv: Tensor, wo: Tensor, bo: Optional[Tensor],
          heads: int, explicit: bool = False) -> Tensor:
    """Whole attn2 module: to_q / to_k / to_v (bias-free) -> core -> to_out[0] (+bias); Dropout(0) is identity.
    Residual add and norm2 are outside the processor (BasicTransformerBlock)."""
    q = F.linear(hidden, wq)
    k = F.linear(ctx, wk)
    v = F.linear(ctx, wv)
    return F.linear(attn2_core(q, k, v, heads, explicit=explicit), wo, bo)


def softmax_entropy(q: Tensor, k: Tensor, heads: int) -> Tensor:
    """Mean entropy (nats) of the attention rows -- recorded by tests so inputs are known not to be trivially
    uniform (SURVEY.md section 8d, config 2 note)."""
    b, n, c = q.shape
    d = c // heads
    qh = q.view(b, n, heads, d).transpose(1, 2).float()
    kh = k.view(b, -1, heads, d).transpose(1, 2).float()
    p = torch.softmax(torch.matmul(qh, kh.transpose(-1, -2)) * (d ** -0.5), dim=-1)
    return -(p * torch.log(p.clamp_min(1e-30))).sum(-1).mean()


# =============================================================================
# A8 / A9 -- base + output-level + feature-level KD loss
# =============================================================================
def _masked_mse_mean(a: Tensor, b: Tensor, w: Tensor) -> Tensor:
    """mean_b( mean_chw( (a-b)^2 * w_b ) ) -- the shape every term of train_sdxl_zh.py:400-441 has."""
    l = F.mse_loss(a.float(), b.float(), reduction="none") * w.view(-1, 1, 1, 1).float()
    return l.mean(dim=[1, 2, 3]).mean()


def kd_loss(noise_pred: Tensor, noise: Tensor, noise_pred_teacher: Tensor,
            feats_student: Sequence[Tensor], feats_teacher: Sequence[Tensor], zh_or_not: Tensor,
            feature_weight: float = 0.1, skip_nonfinite: bool = False) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """Total loss and its three parts (train_sdxl_zh.py:400-441; NaN/Inf tap skipping of
    train_sd_zh.py:246-268 when ``skip_nonfinite``).

      L0 = mean_b(mean_chw((eps_S - eps)^2 * zh_b))
      L1 = mean_b(mean_chw((eps_S - eps_T)^2 * (1 - zh_b)))
      L2 = sum_taps mean_b(mean_chw((f_T - f_S)^2 * (1 - zh_b)))
      L  = L0 + L1 + feature_weight * L2

    ``zh_or_not`` is the int64 [B] language flag of the batch; gradients flow to the student tensors only.
    """
    zh = zh_or_not.float()
    l0 = _masked_mse_mean(noise_pred, noise, zh)
    l1 = _masked_mse_mean(noise_pred, noise_pred_teacher.detach(), 1.0 - zh)
    l2 = noise_pred.new_zeros((), dtype=torch.float32)
    for fs, ft in zip(feats_student, feats_teacher):
        term_full = F.mse_loss(ft.detach().float(), fs.float(), reduction="none") * (1.0 - zh).view(-1, 1, 1, 1)
        if skip_nonfinite and not bool(torch.isfinite(term_full).all()):
            continue
        l2 = l2 + term_full.mean(dim=[1, 2, 3]).mean()
    return l0 + l1 + feature_weight * l2, l0, l1, l2


# =============================================================================
# KD taps (A7) and the whole step (configs 1-3), hosted on the UNet shell
# =============================================================================
def install_taps(unet, store: Dict[str, Tensor]) -> list:
    """train_sdxl_zh.py:69-84: down blocks -> output[0]; mid / up blocks -> output.  Stores references, no copy."""
    handles = []
    for i, blk in enumerate(unet.down_blocks):
        handles.append(blk.register_forward_hook(lambda m, a, out, k=f"d{i}": store.__setitem__(k, out[0])))
    handles.append(unet.mid_block.register_forward_hook(lambda m, a, out: store.__setitem__("m", out)))
    for i, blk in enumerate(unet.up_blocks):
        handles.append(blk.register_forward_hook(lambda m, a, out, k=f"u{i}": store.__setitem__(k, out)))
    return handles


def tap_order(store: Mapping[str, Tensor]) -> List[str]:
    """d0..d{n-1}, m, u0..u{n-1} -- the summation order of train_sdxl_zh.py:424-438."""
    n = sum(1 for k in store if k.startswith("d"))
    return [f"d{i}" for i in range(n)] + ["m"] + [f"u{i}" for i in range(n)]


def kd_train_step(unet_student, unet_teacher, adapter_sd_: Mapping[str, Tensor], batch: Mapping[str, Tensor],
                  variant: str = "sdxl", skip_nonfinite: Optional[bool] = None):
    """One KD training step's forward graph on host tensors (train_sdxl_zh.py:383-441 / train_sd_zh.py:207-272).

    ``batch`` holds synthetic tensors of the reference's shapes: noisy_latents, noise,
    timesteps, enc, enc_uncond (non-English CLIP states), teacher_ctx, teacher_ctx_uncond, [teacher_pooled,
    time_ids], zh_or_not, prompt_mask.  Returns (loss, l0, l1, l2); call ``loss.backward()`` for adapter grads.
    """
    if skip_nonfinite is None:
        skip_nonfinite = variant == "sd"
    taps_s: Dict[str, Tensor] = {}
    taps_t: Dict[str, Tensor] = {}
    hs = install_taps(unet_student, taps_s) + install_taps(unet_teacher, taps_t)
    try:
        mask = batch["prompt_mask"]
        if variant == "sdxl":
            x1, ctx = adapter_sdxl(adapter_sd_, batch["enc"])
            _x1u, ctx_u = adapter_sdxl(adapter_sd_, batch["enc_uncond"])
            ctx = cfg_select(mask, ctx_u, ctx)
            kw_s = {"text_embeds": x1, "time_ids": batch["time_ids"]}
            kw_t = {"text_embeds": batch["teacher_pooled"], "time_ids": batch["time_ids"]}
        else:
            ctx = cfg_select(mask, adapter_sd(adapter_sd_, batch["enc_uncond"]), adapter_sd(adapter_sd_, batch["enc"]))
            kw_s = kw_t = None
        pred = unet_student(batch["noisy_latents"], batch["timesteps"], ctx, added_cond_kwargs=kw_s)[0]
        with torch.no_grad():
            t_ctx = cfg_select(mask, batch["teacher_ctx_uncond"], batch["teacher_ctx"])
            pred_t = unet_teacher(batch["noisy_latents"], batch["timesteps"], t_ctx, added_cond_kwargs=kw_t)[0]
        order = tap_order(taps_s)
        return kd_loss(pred, batch["noise"], pred_t, [taps_s[k] for k in order], [taps_t[k] for k in order],
                       batch["zh_or_not"], skip_nonfinite=skip_nonfinite)
    finally:
        for h in hs:
            h.remove()
