// common.cuh -- shared host/device helpers for libpea_b200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/pea_b200.h"

namespace pea {

// ---- host-side error plumbing --------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define PEA_CHECK_ARG(cond, ...)                \
  do {                                          \
    if (!(cond)) {                              \
      pea::set_error(__VA_ARGS__);              \
      return PEA_EINVAL;                        \
    }                                           \
  } while (0)

#define PEA_CUDA(expr)                                                                       \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      pea::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return PEA_ECUDA;                                                                      \
    }                                                                                        \
  } while (0)

#define PEA_LAUNCH_CHECK()                                                                   \
  do {                                                                                       \
    cudaError_t _e = cudaGetLastError();                                                     \
    if (_e != cudaSuccess) {                                                                 \
      pea::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return PEA_ECUDA;                                                                      \
    }                                                                                        \
    pea::count_launch();                                                                     \
  } while (0)

inline size_t dtype_size(pea_dtype t) { return t == PEA_F32 ? 4 : 2; }
int num_sms();

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device).  The attribute belongs to the device's copy
// of the function, and although the design is one process per GPU nothing here should break when one process drives
// several: `done` is a per-kernel bit mask indexed by device ordinal.
template <typename Kernel>
inline int ensure_dynamic_smem(Kernel kernel, int bytes, std::atomic<unsigned long long>& done) {
  int dev = 0;
  PEA_CUDA(cudaGetDevice(&dev));
  const unsigned long long bit = 1ull << (dev & 63);
  if (!(done.load(std::memory_order_acquire) & bit)) {
    PEA_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    done.fetch_or(bit, std::memory_order_release);
  }
  return PEA_OK;
}

// Launch through cudaLaunchKernelEx; `pdl` adds cudaLaunchAttributeProgrammaticStreamSerialization (the kernel must then
// order its global-memory accesses with pdl_wait(), below).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_ex(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_ex(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, bool pdl, Args&&... args) {
  return launch_ex(kernel, dim3((unsigned)grid), dim3((unsigned)block), smem, st, pdl, static_cast<Args&&>(args)...);
}

#define PEA_LAUNCH_EX(...)                 \
  do {                                     \
    PEA_CUDA(launch_ex(__VA_ARGS__));      \
    PEA_LAUNCH_CHECK();                    \
  } while (0)

// ---- device helpers ------------------------------------------------------------------------------------------
// Programmatic dependent launch.  pdl_wait(): block until the grid this one was launched behind has completed and its
// memory is visible (returns at once for an ordinary launch).  pdl_launch_dependents(): allow the NEXT grid in the stream,
// if it was launched with the PDL attribute, to start its prologue as SM resources free up.  Issued in this order, a
// running grid implies that everything before its predecessor has completed.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename T> struct VecTraits;

__device__ __forceinline__ float to_float(float v) { return v; }
__device__ __forceinline__ float to_float(__half v) { return __half2float(v); }
__device__ __forceinline__ float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ float from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_float<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 16-bit pair packing for MMA fragments
template <typename T> __device__ __forceinline__ uint32_t pack2(float lo, float hi);
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <> __device__ __forceinline__ uint32_t pack2<__half>(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <typename T> __device__ __forceinline__ float2 unpack2(uint32_t u);
template <> __device__ __forceinline__ float2 unpack2<__nv_bfloat16>(uint32_t u) {
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
}
template <> __device__ __forceinline__ float2 unpack2<__half>(uint32_t u) {
  return __half22float2(*reinterpret_cast<__half2*>(&u));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// exact-erf GELU and its derivative (nn.GELU() default, train_sdxl_zh.py:51,53,62)
__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_exact_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// n / d for 0 <= n < 2^23, d >= 1 without the ~300-cycle emulated integer division: float quotient, then exact fix-up.
// (Two of these at kernel entry replaced ~650 cycles of every warp's prologue in the attention kernels.)
__device__ __forceinline__ int fast_div(int n, int d) {
  int q = __float2int_rz(__fdividef((float)n, (float)d));
  q -= (q * d > n);
  q += ((q + 1) * d <= n);
  return q;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// streaming 128-bit global accesses (read-once / write-once data: keep it out of L1)
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

}  // namespace pea
