#!/usr/bin/env python
"""Which ingredient breaks a SECOND capture of the step (bench.py's measurement graphs)?  usage: capture_probe.py overlap timer warm"""
import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pea_diffusion_b200 import ops
from pea_diffusion_b200.train import KDTrainer
from tests.step_util import build_product, device_batch, make_case
overlap, timer, warm = (int(x) for x in sys.argv[1:4])
DEV = "cuda:0"
case = make_case("sdxl", tiny=True, batch=2, seed=5)
s, t, m = build_product(case, DEV, torch.bfloat16)
tr = KDTrainer(s, t, m, "sdxl", lr=1e-3, cuda_graph=True, overlap_teacher=bool(overlap))
b = device_batch(case, DEV, torch.bfloat16)
for _ in range(3):
    tr.step(b)
torch.cuda.synchronize()
ts, tr.teacher_stream = tr.teacher_stream, None
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
if warm:
    with torch.cuda.stream(side):
        tr._forward_backward(tr._static_in)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
try:
    if timer:
        ops.TIMER.reset(True, external=True)
    with torch.cuda.graph(g, stream=side):
        tr._forward_backward(tr._static_in)
    ops.TIMER.enabled = False
    g.replay(); torch.cuda.synchronize()
    print("PROBE overlap=%d timer=%d warm=%d OK" % (overlap, timer, warm), ops.TIMER.summary().keys() if timer else "")
except Exception as e:
    print("PROBE overlap=%d timer=%d warm=%d FAIL %s" % (overlap, timer, warm, str(e).splitlines()[0][:120]))
