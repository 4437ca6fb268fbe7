#!/usr/bin/env python
"""What one attn2 launch costs INSIDE a CUDA graph, without event brackets: replay a graph of R x [to_q GEMM -> attn2 core
(-> to_out GEMM)] and the same graph without the attention launches; the difference / R is the marginal device time of one
attention launch with its input freshly written by the producer GEMM (L2-resident), i.e. the situation of a training step.
Bracketed figures (event-record nodes around the launch, as bench.py's per-kernel table uses) are printed beside it.

    python tools/chain_probe.py [--mode fwd|bwd] [--N 1024] [--C 1280] [--reps 20] [--own-gemm 0|1]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pea_diffusion_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="fwd")
ap.add_argument("--N", type=int, default=1024)
ap.add_argument("--C", type=int, default=1280)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--own-gemm", type=int, default=0, help="1: the producer / consumer GEMMs are pea_gemm (PDL chain when built in)")
ap.add_argument("--replays", type=int, default=20)
ap.add_argument("--pdl", type=int, default=0, help="1: every library launch carries PEA_LAUNCH_PDL")
args = ap.parse_args()

dev = "cuda:0"
B, L, N, Cc = 8, 77, args.N, args.C
H = Cc // 64
dt = torch.bfloat16
h = torch.randn(B * N, Cc, device=dev, dtype=dt)
wq = torch.randn(Cc, Cc, device=dev, dtype=dt) * Cc ** -0.5
wo = torch.randn(Cc, Cc, device=dev, dtype=dt) * Cc ** -0.5
kv = torch.randn(B, L, 2 * Cc, device=dev, dtype=dt)
q = torch.empty(B * N, Cc, device=dev, dtype=dt)
o2 = torch.empty(B * N, Cc, device=dev, dtype=dt)
do = torch.randn(B, N, Cc, device=dev, dtype=dt)
dkv = torch.zeros(B, L, 2 * Cc, device=dev, dtype=torch.float32)


def gemm(a, w, out):
    if args.own_gemm:
        ops.gemm([a], [w], a.shape[0], w.shape[0], a.shape[1], out0=out, pdl=bool(args.pdl))
    else:
        torch.matmul(a, w.t(), out=out)


def body(with_attn, bracket=None):
    for _ in range(args.reps):
        gemm(h, wq, q)                                   # producer: writes Q (stays in L2)
        if with_attn:
            if bracket is not None:
                e0 = torch.cuda.Event(enable_timing=True, external=True); e0.record()
            if args.mode == "fwd":
                o = ops.attn2_fwd(q.view(B, N, Cc), kv, H, 0.125, pdl=bool(args.pdl))
            else:
                o, _ = ops.attn2_bwd(q.view(B, N, Cc), kv, do, H, 0.125, dkv=dkv, pdl=bool(args.pdl))
            if bracket is not None:
                e1 = torch.cuda.Event(enable_timing=True, external=True); e1.record()
                bracket.append((e0, e1))
        else:
            o = do
        gemm(o.view(B * N, Cc), wo, o2)                   # consumer: reads O / dQ


def capture(with_attn, bracket=None):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        body(with_attn)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        body(with_attn, bracket)
    return g


def time_graph(g):
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(args.replays):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


g_with = capture(True)
g_without = capture(False)
pairs = []
g_br = capture(True, pairs)
t_with, t_without = time_graph(g_with), time_graph(g_without)
g_br.replay(); g_br.replay(); torch.cuda.synchronize()
br = sorted(e0.elapsed_time(e1) * 1e3 for e0, e1 in pairs)
es = 2
nbytes = (2 * B * N * Cc * es + 2 * B * L * Cc * es) if args.mode == "fwd" else (3 * B * N * Cc * es + 4 * B * L * Cc * es)
marg = (t_with - t_without) / args.reps
print(json.dumps({"probe": "chain", "mode": args.mode, "N": N, "C": Cc, "own_gemm": args.own_gemm, "pdl": args.pdl, "lib": os.environ.get("PEA_B200_LIB", "product"),
                  "graph_with_us": t_with, "graph_without_us": t_without, "marginal_us_per_launch": marg,
                  "bracketed_us_median": br[len(br) // 2], "algorithmic_bytes": nbytes,
                  "marginal_GBps": nbytes / marg / 1e3, "bracketed_GBps": nbytes / br[len(br) // 2] / 1e3}))
