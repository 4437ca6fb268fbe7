#!/usr/bin/env python
"""A few dozen launches of every kernel family on ragged shapes, meant to run under compute-sanitizer:

    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py

Results are not checked here (tests/ does that); the point is out-of-bounds / misaligned accesses in the epilogues, the
atomics and the TMA boxes at the edges.  Summaries: profiles/r1_sanitizer.txt, profiles/r2_sanitizer.txt.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pea_diffusion_b200 import ops
from pea_diffusion_b200.kd_loss import kd_loss

dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s, dt=torch.bfloat16: torch.randn(*s, device=dev, generator=g).to(dt)

# GEMM: ragged M/N/K, both majors, split planes, fused epilogues, split-K accumulate
for (M, N, K) in [(77, 72, 40), (130, 200, 136), (616, 1280, 2048)]:
    for bk in (True, False):
        A, B = [rn(M, K)], [rn(N, K) if bk else rn(K, N)]
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        ops.gemm(A, B, M, N, K, b_kmajor=bk, out0=out)
    A2, B2 = [rn(M, K), rn(M, K)], [rn(N, K), rn(N, K)]
    o32 = torch.zeros(M, N, device=dev, dtype=torch.float32)
    bias = rn(N, dt=torch.float32)
    planes = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(2)]
    ops.gemm(A2, B2, M, N, K, bias=bias, act=True, out0=o32, out1=planes)
    ops.gemm(A2[:1], B2[:1], M, N, K, out0=o32, accumulate=True, split_k=2)
M, N, K = 4096, 512, 256   # CTA-pair path
ops.gemm([rn(M, K)], [rn(N, K)], M, N, K, out0=torch.empty(M, N, device=dev, dtype=torch.bfloat16))

# attn2: tcgen05 path (d = 64) with ragged N / L and several tiles per CTA; mma.sync path (d = 40, 160)
for (B, N, L, H, d) in [(2, 1000, 52, 10, 64), (40, 100, 77, 20, 64), (1, 130, 80, 2, 64), (3, 16, 1, 4, 64), (2, 70, 77, 8, 40), (2, 16, 77, 8, 160)]:
    for dt in (torch.bfloat16, torch.float16):
        C = H * d
        q, kv, do = rn(B, N, C, dt=dt), rn(B, L, 2 * C, dt=dt), rn(B, N, C, dt=dt)
        ops.attn2_fwd(q, kv, H, d ** -0.5)
        ops.attn2_bwd(q, kv, do, H, d ** -0.5)

# programmatic dependent launch chain (round 2): to_q GEMM -> attn2 -> to_out GEMM with bias, all PEA_LAUNCH_PDL, and its backward
B, N, L, H, d = 2, 300, 77, 4, 64
C = H * d
h, wq, wo, kv = rn(B * N, C), rn(C, C), rn(C, C), rn(B, L, 2 * C)
bias = rn(C, dt=torch.float32)
q = torch.empty(B * N, C, device=dev, dtype=torch.bfloat16)
out = torch.empty(B * N, C, device=dev, dtype=torch.bfloat16)
for _ in range(2):
    ops.gemm([h], [wq], B * N, C, C, out0=q, pdl=True)
    o = ops.attn2_fwd(q.view(B, N, C), kv, H, d ** -0.5, pdl=True)
    ops.gemm([o.view(B * N, C)], [wo], B * N, C, C, bias=bias, out0=out, pdl=True)
    ops.gemm([out], [wo], B * N, C, C, b_kmajor=False, out0=q, pdl=True)
    dq, _ = ops.attn2_bwd(h.view(B, N, C), kv, q.view(B, N, C), H, d ** -0.5, pdl=True)

# KD loss (ragged tap sizes), row ops
shp = [(2, 4, 9, 9), (2, 320, 7, 5), (2, 1280, 3, 3)]
fs = [rn(*s).requires_grad_(True) for s in shp[1:]]
ft = [rn(*s) for s in shp[1:]]
pred = rn(*shp[0]).requires_grad_(True)
loss = kd_loss(pred, rn(*shp[0]), rn(*shp[0]), fs, ft, torch.tensor([0, 1], device=dev))[0]
loss.backward()
x = rn(154, 1024, dt=torch.float32)
pl, mean, rstd = ops.layernorm_fwd(x, torch.ones(1024, device=dev), torch.zeros(1024, device=dev), 1e-5, 2, torch.bfloat16)
ops.split_planes(rn(77, 1000, dt=torch.float32)[:, :996], 3, torch.bfloat16)
ops.cfg_select(torch.tensor([1, 0], dtype=torch.uint8, device=dev), rn(2, 77, 2048), rn(2, 77, 2048))
# the adapter module: its ~30 launches (LayerNorm, plane splits, GEMMs, column sums, mean-pool) are one PDL chain
from pea_diffusion_b200 import adapter as A
for prec in ("fp32", "bf16"):
    m = A.MLP(128, 160, 128, 256, False, precision=prec).to(dev)
    xa = rn(3, 77, 128, dt=torch.float32).requires_grad_(True)
    x1, x2 = m(xa)
    (x1.square().sum() + x2.square().sum()).backward()
torch.cuda.synchronize()
print("sanitize_smoke: all launches completed")
