#!/usr/bin/env python
"""Per-tile event timeline of attn2_bwd_tc_kernel (CTA 0) from an instrumented build (-DPEA_TRACE):
    sh tools/build_trace.sh && PEA_B200_LIB=build/libpea_trace.so python tools/trace_bwd.py [N] [C]
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pea_diffusion_b200 import _lib, ops

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
Cc = int(sys.argv[2]) if len(sys.argv) > 2 else 1280
H, B, L = Cc // 64, 8, 77
q = torch.randn(B, N, Cc, device="cuda", dtype=torch.bfloat16)
kv = torch.randn(B, L, 2 * Cc, device="cuda", dtype=torch.bfloat16)
do = torch.randn(B, N, Cc, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.attn2_bwd(q, kv, do, H, 0.125)
torch.cuda.synchronize()
n = 16 * 64 * 16
buf = (C.c_ulonglong * n)()
dll = C.CDLL(_lib.LIB_PATH)
assert dll.pea_debug_bwd_trace(buf, n) == 0
ev = lambda role, t, e: buf[(role * 64 + t) * 16 + e]
t0 = ev(0, 63, 15)
print("kernel entry -> all roles done:", ev(0, 63, 14) - t0, "cycles; groups done at", ev(2, 63, 14) - t0, ev(3, 63, 14) - t0)
names = {0: ["ring_empty"], 1: ["sdp:enter", "sdp:ring_full", "main:enter", "ds_ready", "dq_free"],
         2: ["enter", "sdp_full", "pre_pds", "pds_free", "ds_arrive"]}
ntiles = (B * H * ((N + 127) // 128) + 147) // 148
lo, hi = (8, 20) if ntiles > 24 else (0, ntiles)
print(f"cycles relative to kernel entry; tiles {lo}..{hi - 1} of {ntiles}")
for t in range(lo, hi):
    g = t & 1
    row = [f"t{t:2d} g{g}"]
    row.append("prod " + " ".join(f"{ev(0, t, i) - t0:7d}" for i in range(1)))
    row.append("| mma " + " ".join(f"{names[1][i]}={ev(1, t, i) - t0:7d}" for i in range(5)))
    row.append("| sm " + " ".join(f"{names[2][i]}={ev(2 + g, t, i) - t0:7d}" for i in range(5)))
    row.append(f"| dq_epilogue_done={ev(4, t, 0) - t0:7d}")
    print(" ".join(row))
print("softmax passes A / B / C (cycles):", [(ev(2 + (t & 1), t, 8) - ev(2 + (t & 1), t, 1), ev(2 + (t & 1), t, 9) - ev(2 + (t & 1), t, 8),
                                               ev(2 + (t & 1), t, 2) - ev(2 + (t & 1), t, 9)) for t in range(lo, hi)])
print("per-tile period (MMA ds_ready to ds_ready):", [ev(1, t + 1, 3) - ev(1, t, 3) for t in range(lo, hi - 1)])
