#!/usr/bin/env python
"""BASELINE.json configs[4]: the forward-only attn2 path of tests/test_sdxl_zh.py -- SDXL-shaped UNet, 1024 px, fp16,
UNet batch 8 (4 prompts x classifier-free guidance), 30 denoise iterations with a CONSTANT context (so every layer's
K|V projection is loop invariant: tests/test_sdxl_zh.py:378-391), scheduler math replaced by `latents -= 0.01 * eps`.

Prints one JSON line per arm: stock processors (what the reference runs), PEAAttn2Processor, PEAAttn2Processor with the
K|V cache.  Context only -- bench.py is the contract benchmark.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pea_diffusion_b200 import _lib, attn2
from pea_diffusion_b200.shell.unet import StockAttnProcessor, UNet2DConditionShell, UNetConfig

DEV = "cuda:0"


def main(steps=30):
    cfg = UNetConfig.sdxl()
    torch.manual_seed(0)
    with torch.device(DEV):
        unet = UNet2DConditionShell(cfg)
    unet = unet.half().requires_grad_(False)
    g = torch.Generator(device=DEV).manual_seed(1)
    ctx = torch.randn(8, 77, 2048, device=DEV, dtype=torch.float16, generator=g)
    kw = {"text_embeds": torch.randn(8, 1280, device=DEV, dtype=torch.float16, generator=g),
          "time_ids": torch.tensor([[1024.0, 1024, 0, 0, 1024, 1024]] * 8, device=DEV)}
    lat0 = torch.randn(8, 4, 128, 128, device=DEV, dtype=torch.float16, generator=g)
    outs = {}
    for arm in ("stock", "pea", "pea_cache_kv"):
        if arm == "stock":
            unet.set_attn_processor(StockAttnProcessor())
        else:
            unet.set_attn_processor(StockAttnProcessor())
            attn2.install(unet, cache_kv=(arm == "pea_cache_kv"))

        def loop():
            lat = lat0.clone()
            with torch.no_grad():
                for i in range(steps):
                    eps = unet(lat, torch.tensor([999 - 33 * i], device=DEV), ctx, added_cond_kwargs=kw)[0]
                    lat = lat - 0.01 * eps
            return lat

        loop()  # warm-up (also fills the K|V cache: the first pass of a real run pays it once)
        torch.cuda.synchronize()
        n0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        outs[arm] = loop()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(json.dumps({"arm": arm, "denoise_steps": steps, "unet_batch": 8, "ms_total": ms, "ms_per_unet_call": ms / steps,
                          "images_per_s": 4 / (ms * 1e-3), "pea_kernel_launches": _lib.launch_count() - n0}), flush=True)
    ref = outs["stock"].float()
    for arm in ("pea", "pea_cache_kv"):
        err = float((outs[arm].float() - ref).norm() / ref.norm())
        print(json.dumps({"arm": arm, "rel_err_vs_stock_after_%d_steps" % steps: err}), flush=True)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 30)
