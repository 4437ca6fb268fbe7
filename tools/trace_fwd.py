#!/usr/bin/env python
"""Per-tile event timeline of attn2_fwd_tc_kernel (CTA 0) from an instrumented build (-DPEA_TRACE):
    PEA_B200_LIB=build/libpea_trace.so python tools/trace_fwd.py [N] [C]
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pea_diffusion_b200 import _lib, ops

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
Cc = int(sys.argv[2]) if len(sys.argv) > 2 else 1280
H, B, L = Cc // 64, 8, 77
q = torch.randn(B, N, Cc, device="cuda", dtype=torch.bfloat16)
kv = torch.randn(B, L, 2 * Cc, device="cuda", dtype=torch.bfloat16)
scratch = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
WARM = len(sys.argv) > 3 and sys.argv[3] == "warm"   # Q freshly written (L2-resident), as inside a training step
src = torch.randn(B, N, Cc, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    scratch.view(torch.int32).sum()
    if WARM:
        q.copy_(src)
    ops.attn2_fwd(q, kv, H, 0.125)
torch.cuda.synchronize()
n = 16 * 64 * 16
buf = (C.c_ulonglong * n)()
dll = C.CDLL(_lib.LIB_PATH)
assert dll.pea_debug_fwd_trace(buf, n) == 0
ev = lambda role, t, e: buf[(role * 64 + t) * 16 + e]
t0 = ev(0, 0, 15)
print(f"N={N} C={Cc}: cycles since kernel entry of CTA 0 ({'Q in L2' if WARM else 'cold L2'})")
print("prologue done", ev(0, 0, 14) - t0, "| first K|V TMA issued", ev(0, 0, 13) - t0, "| K|V landed", ev(1, 0, 5) - t0,
      "| Q0 landed", ev(1, 0, 1) - t0, "| all roles done", ev(0, 0, 12) - t0)
print("per-warp prologue (entry / own work done / past the barrier):",
      " ".join(f"w{w}:{ev(15, w, 0) - t0}/{ev(15, w, 1) - t0}/{ev(15, w, 2) - t0}" for w in range(15)))
ntiles = (B * H * ((N + 127) // 128) + 147) // 148
for t in range(min(ntiles, 12)):
    g = t % 3   # three softmax groups, tiles dealt round-robin (roles 2, 3, 4 in the trace buffer)
    print(f"t{t:2d} g{g} prod issue {ev(0, t, 0) - t0:6d} | S-warp q_wait {ev(1, t, 0) - t0:6d} q_full {ev(1, t, 1) - t0:6d} | PV-warp p_wait {ev(1, t, 2) - t0:6d} p_ready {ev(1, t, 3) - t0:6d}"
          f" | sm s_wait {ev(2 + g, t, 0) - t0:6d} s_full {ev(2 + g, t, 1) - t0:6d} p_arrive {ev(2 + g, t, 2) - t0:6d} o_wait {ev(2 + g, t, 3) - t0:6d} o_full {ev(2 + g, t, 4) - t0:6d} stored {ev(2 + g, t, 5) - t0:6d}")
print("group done:", ev(2, 63, 6) - t0, ev(3, 63, 6) - t0, ev(4, 63, 6) - t0)
