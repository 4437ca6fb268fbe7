#!/usr/bin/env python
"""Tiny KD trainer with to_q / to_out on the library GEMM as a PDL chain (own_projections=True): eager, CUDA graph, and CUDA
graph with the teacher on a parallel branch must produce the loss trajectory of the default (cuBLAS projections) trainer.
Run under `timeout`: a programmatic-dependency deadlock would otherwise hold the GPU box."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pea_diffusion_b200 import attn2
from pea_diffusion_b200.train import KDTrainer
from tests.step_util import _unets, device_batch, make_case
from pea_diffusion_b200 import adapter as A

DEV = "cuda:0"
case = make_case("sdxl", tiny=True, batch=2, seed=5)
out = {}
for own, graph, overlap in ((False, False, False), (True, False, False), (True, True, False), (True, True, True)):
    s, t = _unets(case, DEV, torch.bfloat16)
    attn2.install(s, own_projections=own)
    attn2.install(t, own_projections=own)
    sd = case["adapter_sd"]
    m = A.MLP(case["dims"][0], sd["projector.4.weight"].shape[0], case["dims"][1], sd["fc.weight"].shape[0], False)
    m.load_state_dict(sd)
    m = m.to(DEV)
    tr = KDTrainer(s, t, m, "sdxl", lr=1e-3, cuda_graph=graph, overlap_teacher=overlap)
    b = device_batch(case, DEV, torch.bfloat16)
    losses = [float(tr.step(b)[0]) for _ in range(4)]
    torch.cuda.synchronize()
    out[(own, graph, overlap)] = losses
    print(own, graph, overlap, [round(x, 5) for x in losses], flush=True)
    tr.close()
ref = out[(False, False, False)]
for k, v in out.items():
    assert all(abs(a - b) <= 2e-2 * abs(b) for a, b in zip(v, ref)), (k, v, ref)
print("pdl_probe ok")
