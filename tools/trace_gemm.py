#!/usr/bin/env python
"""Event timeline of gemm_tc_kernel (CTA 0) from an instrumented build (-DPEA_TRACE):
    PEA_B200_LIB=build/variants/libpea_trace.so python tools/trace_gemm.py M N K [out_dtype: bf16|f32]
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pea_diffusion_b200 import _lib, ops

M, N, K = (int(x) for x in sys.argv[1:4])
odt = torch.float32 if (len(sys.argv) > 4 and sys.argv[4] == "f32") else torch.bfloat16
A = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
B = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
o = torch.empty(M, N, device="cuda", dtype=odt)
scratch = torch.empty(64 << 20, device="cuda", dtype=torch.int32)
for _ in range(3):
    scratch.sum()
    ops.gemm([A], [B], M, N, K, out0=o)
torch.cuda.synchronize()
buf = (C.c_ulonglong * 64)()
dll = C.CDLL(_lib.LIB_PATH)
assert dll.pea_debug_gemm_trace(buf, 64) == 0
names = ["entry", "prologue done", "first stage landed", "last stage of tile 0 landed", "acc 0 complete", "epilogue 0 done", "all roles done"]
print(f"gemm {M}x{N}x{K} -> {odt}: cycles since kernel entry of CTA 0 (cold L2)")
for i, n in enumerate(names):
    print(f"  {n:30s} {buf[i] - buf[0]:8d}")
