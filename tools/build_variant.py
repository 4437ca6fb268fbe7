#!/usr/bin/env python
"""Build an instrumented / experimental copy of libpea_b200.so next to the product library (same sources, extra -D flags):

    python tools/build_variant.py trace -DPEA_TRACE            -> build/variants/libpea_trace.so
    PEA_B200_LIB=build/variants/libpea_trace.so python tools/trace_fwd.py

Used for same-box A/B runs of kernel changes (box-to-box spread is +-5 %, so a change is only accepted or rejected on two
builds timed in ONE gpurun call) and for the -DPEA_TRACE clock64() timelines.  `build/` is git-ignored but travels to the
GPU box with the snapshot.
"""
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as G  # noqa: E402


def build_variant(name, defines):
    objdir = os.path.join(ROOT, "build", "variants", name)
    os.makedirs(objdir, exist_ok=True)
    lib = os.path.join(ROOT, "build", "variants", f"libpea_{name}.so")
    nvcc = G._nvcc()
    srcs = sorted(glob.glob(os.path.join(G.CSRC, "*.cu")))

    def one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        subprocess.run([nvcc, *G.NVCC_FLAGS, *defines, "-c", src, "-o", obj], check=True, cwd=G.CSRC)
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(one, srcs))
    subprocess.run([nvcc, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a"], check=True)
    return lib


if __name__ == "__main__":
    print(build_variant(sys.argv[1], sys.argv[2:]))
