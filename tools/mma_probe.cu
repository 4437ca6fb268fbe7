// mma_probe.cu -- how long does the tensor pipe take per tcgen05.mma at the SMALL shapes the attn2 kernels use?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I pea_diffusion_b200/csrc -I include \
//        tools/mma_probe.cu -o build/mma_probe && build/mma_probe
//
// One CTA per SM (all 148, so shared-memory / tensor pipes are loaded as in the real kernels is NOT the point -- each CTA is
// independent), one thread issues R MMAs of one configuration back to back, commits, and the cycle count from first issue to
// the commit barrier completing is divided by R.  Also measured: the cycles the ISSUING thread needs to get through the R
// issue instructions (does issue block?).  Operand data is whatever is in shared memory / TMEM: timing does not depend on it.
// Results of the round-1 run: profiles/r1_mma_probe.txt.
#include <cstdio>
#include <cuda_runtime.h>
#include "tc05.cuh"

using namespace pea;

struct Cfg { int ts; int M; int N; int a_mn; int b_mn; int chains; int reps; };
constexpr int kReps = 64;

// Everything the issue loop touches is a compile-time constant (a first version with runtime configuration measured its own
// scalar loop: 200 cycles per iteration whatever the shape).
template <int TS, int A_MN, int B_MN, int CHAINS>
__global__ void __launch_bounds__(128, 1) probe(uint32_t idesc, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { tc::mbar_init(smem_u32(&bar), 1); tc::fence_mbar_init(); }
  if (warp == 1) tc::tmem_alloc(smem_u32(&tmem_slot), 512);
  for (uint32_t i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u;
  tc::fence_proxy_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    // A at smem+0 (two 16 KB slabs), B at smem+32 KB; descriptors as in the kernels (SW128; K-major: SBO 1024; MN-major: LBO 16 KB)
    constexpr uint64_t hiK = ((uint64_t)1 << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
    constexpr uint64_t hiMN = ((uint64_t)(16384u >> 4) << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
    const uint64_t da = (A_MN ? hiMN : hiK) | (uint64_t)(smem >> 4);
    const uint64_t db = (B_MN ? hiMN : hiK) | (uint64_t)((smem + 32768u) >> 4);
    for (int rep = 0; rep < 3; ++rep) {   // rep 0, 1 warm up; rep 2 is reported
      long long t0 = 0, t1 = 0;
      if (tc::elect_one()) {
        t0 = clock64();
#pragma unroll
        for (int i = 0; i < kReps; ++i) {
          const uint32_t d = tmem + 256u + (uint32_t)((i % CHAINS) * 80);   // accumulators at columns 256.. (<= 3 chains of 80)
          constexpr int dummy = 0; (void)dummy;
          const uint32_t k = (uint32_t)(i & 3);
          if (TS) tc::mma_f16_ts(d, tmem + k * 8, db + (B_MN ? k * 128 : k * 2), idesc, i >= CHAINS);
          else tc::mma_f16_ss(d, da + (A_MN ? k * 128 : k * 2), db + (B_MN ? k * 128 : k * 2), idesc, i >= CHAINS);
        }
        t1 = clock64();
        tc::mma_commit(smem_u32(&bar));
      }
      __syncwarp();
      tc::mbar_wait(smem_u32(&bar), (uint32_t)rep & 1u);
      const long long t2 = clock64();
      if (rep == 2 && blockIdx.x == 0 && t0 != 0) { out[0] = (unsigned long long)(t1 - t0); out[1] = (unsigned long long)(t2 - t0); }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) { tc::fence_after_sync(); tc::tmem_dealloc(tmem, 512); }
}

template <int TS, int A_MN, int B_MN, int CHAINS>
int run(int M, int N, unsigned long long* d_out) {
  cudaMemset(d_out, 0, 16);
  cudaFuncSetAttribute(probe<TS, A_MN, B_MN, CHAINS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
  probe<TS, A_MN, B_MN, CHAINS><<<148, 128, 72 * 1024>>>(tc::make_idesc(1, M, N, A_MN, B_MN), d_out);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned long long h[2] = {0, 0};
  cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
  printf("%s M%3d N%3d A:%s B:%s chains %d | %7.1f | %7.1f %s\n", TS ? "TS" : "SS", M, N, A_MN ? "MN" : "K ", B_MN ? "MN" : "K ", CHAINS,
         (double)h[1] / kReps, (double)h[0] / kReps, e == cudaSuccess ? "" : cudaGetErrorString(e));
  return e == cudaSuccess ? 0 : 1;
}

int main() {
  unsigned long long* d_out;
  cudaMalloc(&d_out, 16);
  printf("# kind shape majors chains | cycles/MMA to completion | cycles/MMA for the issuing thread   (148 CTAs, CTA 0 reported, 64 MMAs)\n");
  int rc = 0;
  rc |= run<0, 0, 0, 1>(128, 80, d_out);    // S / dP of attn2: SS, K-major A and B, one dependent chain
  rc |= run<0, 0, 0, 2>(128, 80, d_out);    //   two interleaved accumulators
  rc |= run<1, 0, 1, 1>(128, 80, d_out);    // P V of the forward: TS (A from TMEM), B = [V|1] MN-major
  rc |= run<1, 0, 1, 2>(128, 80, d_out);
  rc |= run<1, 0, 1, 1>(128, 64, d_out);    // dQ of the backward: TS, N64
  rc |= run<0, 1, 1, 1>(64, 80, d_out);     // dV^T / dK^T of the backward: SS M64, both MN-major
  rc |= run<0, 1, 1, 2>(64, 80, d_out);
  rc |= run<0, 1, 1, 1>(128, 64, d_out);    // the first dV / dK form: SS M128 N64, both MN-major
  rc |= run<0, 0, 0, 1>(128, 128, d_out);   // GEMM BN=128
  rc |= run<0, 0, 0, 1>(128, 256, d_out);   // GEMM BN=256
  rc |= run<0, 0, 0, 1>(128, 16, d_out);    // tiny N: the per-instruction floor
  rc |= run<1, 0, 1, 1>(128, 16, d_out);
  rc |= run<0, 0, 0, 1>(64, 16, d_out);
  return rc;
}
