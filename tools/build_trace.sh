#!/bin/sh
# Instrumented build of the library (-DPEA_TRACE: clock64() event timelines in the attn2 tcgen05 kernels) -> build/libpea_trace.so
set -e
cd "$(dirname "$0")/../pea_diffusion_b200/csrc"
mkdir -p ../../build
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -diag-suppress 177 \
     -DPEA_TRACE -shared -o ../../build/libpea_trace.so *.cu
