"""Shared harness: one KD training step on identical seeded inputs through (a) the product (CUDA kernels behind the
reference-shaped modules) and (b) the CPU oracle.  Used by the GPU parity tests, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / reference arm."""
from __future__ import annotations

import copy
from typing import Dict, Tuple

import torch

from oracle import pea_oracle as O
from pea_diffusion_b200.shell.unet import UNet2DConditionShell, UNetConfig


def make_case(variant: str = "sdxl", tiny: bool = True, batch: int = 2, seed: int = 0, L: int = 77) -> dict:
    """Seeded synthetic weights + batch of the reference's shapes, fp32 on the host."""
    if tiny:
        cfg = UNetConfig.tiny(variant)
        in_dim, hidden = 64, 64
    else:
        cfg = UNetConfig.sdxl() if variant == "sdxl" else UNetConfig.sd15()
        in_dim, hidden = 1024, (1024 if variant == "sdxl" else 2048)
    torch.manual_seed(seed)
    unet = UNet2DConditionShell(cfg)
    g = torch.Generator().manual_seed(1234 + seed)
    hw = cfg.sample_size if tiny else (128 if variant == "sdxl" else 32)
    Dc = cfg.cross_attention_dim
    r = lambda *s: torch.randn(*s, generator=g)
    batch_t = {
        "noisy_latents": r(batch, 4, hw, hw), "noise": r(batch, 4, hw, hw),
        "timesteps": torch.randint(0, 1000, (batch,), generator=g),
        "enc": r(batch, L, in_dim), "enc_uncond": r(batch, L, in_dim),
        "teacher_ctx": r(batch, L, Dc), "teacher_ctx_uncond": r(batch, L, Dc),
        "zh_or_not": torch.tensor([i % 2 for i in range(batch)]),
        "prompt_mask": torch.rand(batch, generator=g) < 0.3,
    }
    if variant == "sdxl":
        pooled = cfg.projection_class_embeddings_input_dim - 6 * cfg.addition_time_embed_dim
        batch_t["teacher_pooled"] = r(batch, pooled)
        batch_t["time_ids"] = torch.tensor([[1024.0, 1024.0, 0.0, 0.0, 1024.0, 1024.0]] * batch)
        adapter_sd = O.adapter_init(in_dim, pooled, hidden, Dc, seed=seed)
    else:
        adapter_sd = O.adapter_init(in_dim, Dc, hidden, None, seed=seed)
    return {"variant": variant, "cfg": cfg, "unet_sd": unet.state_dict(), "adapter_sd": adapter_sd, "batch": batch_t,
            "dims": (in_dim, hidden)}


def _unets(case, device, dtype):
    s = UNet2DConditionShell(case["cfg"])
    s.load_state_dict(case["unet_sd"])
    t = copy.deepcopy(s)  # the reference's teacher is loaded from the same checkpoint (train_sdxl_zh.py:138,151)
    return s.to(device=device, dtype=dtype), t.to(device=device, dtype=dtype)


def run_oracle_step(case) -> Tuple[Tuple[float, float, float, float], Dict[str, torch.Tensor]]:
    """fp32 CPU oracle: stock attention processors + oracle adapter + oracle loss; autograd for the adapter grads."""
    s, t = _unets(case, "cpu", torch.float32)
    sd = {k: v.clone().requires_grad_(True) for k, v in case["adapter_sd"].items()}
    loss, l0, l1, l2 = O.kd_train_step(s, t, sd, case["batch"], variant=case["variant"])
    loss.backward()
    return (float(loss), float(l0), float(l1), float(l2)), {k: v.grad.detach() for k, v in sd.items()}


def build_product(case, device="cuda:0", dtype=torch.bfloat16, adapter_precision=None, fused=True):
    """``fused=False`` keeps the stock processors (torch SDPA -- what the reference runs) in the same shell: the
    comparison that isolates the attn2 kernels from the shell's own bf16 rounding."""
    from pea_diffusion_b200 import adapter as A, attn2
    s, t = _unets(case, device, dtype)
    if fused:
        attn2.install(s, freeze=True)
        attn2.install(t, freeze=True)
    else:
        s.requires_grad_(False)
        t.requires_grad_(False)
    in_dim, hidden = case["dims"]
    sd = case["adapter_sd"]
    if case["variant"] == "sdxl":
        m = A.MLP(in_dim, sd["projector.4.weight"].shape[0], hidden, sd["fc.weight"].shape[0], False, precision=adapter_precision)
    else:
        m = A.MLP_SD(in_dim, sd["projector.4.weight"].shape[0], hidden, precision=adapter_precision)
    m.load_state_dict(sd, strict=True)
    return s, t, m.to(device)


def device_batch(case, device, dtype):
    out = {}
    for k, v in case["batch"].items():
        if k in ("enc", "enc_uncond"):
            out[k] = v.to(device)                      # adapter input stays fp32 (fp32 master weights)
        elif v.is_floating_point():
            out[k] = v.to(device=device, dtype=dtype)
        else:
            out[k] = v.to(device)
    return out


def run_product_step(case, device="cuda:0", dtype=torch.bfloat16, adapter_precision=None, fused=True):
    from pea_diffusion_b200.train import KDTaps, kd_forward
    s, t, m = build_product(case, device, dtype, adapter_precision, fused)
    taps_s, taps_t = KDTaps(s), KDTaps(t)
    loss, l0, l1, l2 = kd_forward(s, t, m, device_batch(case, device, dtype), taps_s, taps_t, case["variant"])
    loss.backward()
    grads = {k: p.grad.detach().float().cpu() for k, p in m.named_parameters()}
    return (float(loss), float(l0), float(l1), float(l2)), grads
