#!/usr/bin/env python
"""Isolated kernel timings at the hot path's real shapes (BASELINE.json configs[3] sweep + the GEMM / loss shapes of
config 2), each launch bracketed by CUDA events, L2 flushed (256 MB scratch write) between iterations.
Library (cuBLAS / SDPA) timings of the same ops are printed beside them for context only.

    python tools/microbench.py [--iters 50] [--out gpurun_out/microbench.jsonl]
"""
import argparse
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from pea_diffusion_b200 import ops

DEV = "cuda:0"
PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))


WARM = 10


def timeit(fn, iters, warm=None):
    scratch = torch.empty(256 << 20, device=DEV, dtype=torch.uint8)
    for _ in range(WARM if warm is None else warm):
        fn()
    ts = []
    for _ in range(iters):
        if os.environ.get("PEA_MB_FLUSH") == "fill":
            scratch.fill_(1)
        else:
            scratch.view(torch.int32).sum()  # L2 flush by READING 256 MB: leaves clean lines (a fill leaves 126 MB of
            #                                    dirty lines whose write-back is then charged to the kernel under test)
        torch.cuda._sleep(600000)  # ~0.3 ms of GPU spin so the host is ahead of the GPU when the timed launch is enqueued
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--out", default="gpurun_out/microbench.jsonl")
    ap.add_argument("--only", default="")
    ap.add_argument("--warm", type=int, default=10)
    ap.add_argument("--filter", default="", help="comma list: gemm shape names / attn 'N:C' pairs to keep")
    args = ap.parse_args()
    global WARM
    WARM = args.warm
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    out = open(args.out, "w")

    def emit(**kw):
        print(json.dumps(kw)); out.write(json.dumps(kw) + "\n"); out.flush()

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import ClockSampler          # nvidia-smi sampler shared with bench.py: clocks / throttle reasons under load
    clocks = ClockSampler(0)
    dt = torch.bfloat16
    keep = [x for x in args.filter.split(",") if x]
    if not args.only or "attn" in args.only:
        B = 8
        for (C, H) in ((640, 10), (1280, 20)):
            for N in (1024, 4096, 16384):
                if keep and f"{N}:{C}" not in keep:
                    continue
                for L in ((77, 64, 52) if (N == 4096 and C == 640 and not keep) else (77,)):
                    q = torch.randn(B, N, C, device=DEV, dtype=dt)
                    kv = torch.randn(B, L, 2 * C, device=DEV, dtype=dt)
                    do = torch.randn(B, N, C, device=DEV, dtype=dt)
                    d = C // H
                    fb, bb = ops.attn2_bytes(B, N, L, C, 2)      # SURVEY.md section 8d formulas
                    tf, _ = timeit(lambda: ops.attn2_fwd(q, kv, H, d ** -0.5), args.iters)
                    tb, _ = timeit(lambda: ops.attn2_bwd(q, kv, do, H, d ** -0.5), args.iters)
                    # library: SDPA on [B,H,N,d] views (what AttnProcessor2_0 runs), fwd and fwd+bwd
                    qh = q.view(B, N, H, d).transpose(1, 2).detach().requires_grad_(True)
                    kh = kv[..., :C].reshape(B, L, H, d).transpose(1, 2).detach().requires_grad_(True)
                    vh = kv[..., C:].reshape(B, L, H, d).transpose(1, 2).detach().requires_grad_(True)
                    doh = do.view(B, N, H, d).transpose(1, 2)
                    with torch.no_grad():
                        ts, _ = timeit(lambda: F.scaled_dot_product_attention(qh, kh, vh), args.iters)
                    o = F.scaled_dot_product_attention(qh, kh, vh)
                    tsb, _ = timeit(lambda: torch.autograd.grad(o, (qh, kh, vh), doh, retain_graph=True), args.iters)
                    emit(op="attn2", B=B, N=N, C=C, H=H, L=L, fwd_us=tf, bwd_us=tb, fwd_GBps=fb / tf / 1e3, bwd_GBps=bb / tb / 1e3,
                         fwd_frac=fb / tf / 1e3 / PEAK["hbm_gbs"], bwd_frac=bb / tb / 1e3 / PEAK["hbm_gbs"],
                         fwdbwd_frac=(fb + bb) / (tf + tb) / 1e3 / PEAK["hbm_gbs"],
                         fwd_TFLOPs=4 * B * N * L * C / tf / 1e6, bwd_TFLOPs=10 * B * N * L * C / tb / 1e6,
                         sdpa_fwd_us=ts, sdpa_bwd_us=tsb)
    if not args.only or "gemm" in args.only:
        shapes = [("kvproj_c640", 616, 1280, 2048, 1, 1, True), ("kvproj_c1280", 616, 2560, 2048, 1, 1, True),
                  ("dctx_c640", 616, 2048, 1280, 2, 1, False), ("dctx_c1280", 616, 2048, 2560, 2, 1, False),
                  ("dctx_c1280_1plane", 616, 2048, 2560, 1, 1, False),
                  ("adapter_fwd_1024", 1232, 1024, 1024, 1, 1, True), ("adapter_fwd_fc", 1232, 2048, 1280, 1, 1, True),
                  ("adapter_fwd_1024_fp32x2", 1232, 1024, 1024, 2, 2, True), ("big_4096", 4096, 4096, 4096, 1, 1, True)]
        for name, M, N, K, sa, sb, bk in shapes:
            if keep and name not in keep:
                continue
            A = [torch.randn(M, K, device=DEV, dtype=dt) for _ in range(sa)]
            Bm = [torch.randn(N, K, device=DEV, dtype=dt) if bk else torch.randn(K, N, device=DEV, dtype=dt) for _ in range(sb)]
            o = torch.empty(M, N, device=DEV, dtype=dt)
            t, tmin = timeit(lambda: ops.gemm(A, Bm, M, N, K, b_kmajor=bk, out0=o), args.iters)
            w = Bm[0] if bk else Bm[0].t()
            tl, _ = timeit(lambda: torch.matmul(A[0], w.t() if bk else Bm[0]), args.iters)
            nprod = sum(1 for i in range(sa) for j in range(sb) if i + j < max(sa, sb))
            emit(op="gemm", name=name, M=M, N=N, K=K, planes=(sa, sb), us=t, min_us=tmin, TFLOPs=2 * M * N * K * nprod / t / 1e6,
                 cublas_us=tl, cublas_TFLOPs=2 * M * N * K / tl / 1e6)
    if not args.only or "kd" in args.only:
        from pea_diffusion_b200.kd_loss import kd_loss
        B = 8
        shp = [(B, 320, 64, 64), (B, 640, 32, 32), (B, 1280, 32, 32), (B, 1280, 32, 32), (B, 1280, 64, 64), (B, 640, 128, 128), (B, 320, 128, 128)]
        fs = [torch.randn(s, device=DEV, dtype=dt).requires_grad_(True) for s in shp]
        ft = [torch.randn(s, device=DEV, dtype=dt) for s in shp]
        pred = torch.randn(B, 4, 128, 128, device=DEV, dtype=dt).requires_grad_(True)
        noise, pt = torch.randn_like(pred), torch.randn_like(pred)
        zh = torch.tensor([0, 1] * 4, device=DEV)
        ops.TIMER.reset(True)
        for _ in range(args.iters):
            kd_loss(pred, noise, pt, fs, ft, zh)
        torch.cuda.synchronize()
        s = ops.TIMER.summary()["kd_mse"]
        ops.TIMER.reset(False)
        emit(op="kd_mse", elems=sum(math.prod(x) for x in shp) + B * 4 * 128 * 128, us=s["ms"] / s["launches"] * 1e3,
             GBps=s["bytes"] / s["ms"] / 1e6, frac=s["bytes"] / s["ms"] / 1e6 / PEAK["hbm_gbs"])
    emit(op="clocks", **clocks.stop())


if __name__ == "__main__":
    main()
