// kd_mse.cu -- fused teacher/student KD loss + gradient (kernels K6+K7).
//
// Replaces the F.mse_loss(reduction="none") * mask -> mean([1,2,3]).mean() chains of
// train_sdxl_zh.py:400-405,416-441 and train_sd_zh.py:218-272 (which make ~5 HBM passes forward and ~4 backward
// over every tap) with ONE pass: read student, read teacher, write d(student).  Algorithmic bytes per element:
// 3 * sizeof(T) (2 * sizeof(T) when no gradient is requested).  HBM-bound; no reuse, so no shared-memory staging:
// 128-bit streaming loads/stores, 32 elements per thread in flight, fp32 accumulation, deterministic two-level
// reduction (per-CTA partials in a workspace, fixed-order finalize -- no float atomics).
#include "common.cuh"

namespace pea {
namespace {

constexpr int kThreads = 256;
constexpr int kElemsPerThread = 32;
constexpr int kChunk = kThreads * kElemsPerThread;  // elements of one sample of one tap handled by one CTA

struct TapParam {
  const void* s;
  const void* t;
  const void* t2;
  void* g;
  const float* w;
  const float* w2;
  long long numel;
  int chunk_begin;        // first CTA index of this tap
  int chunks_per_sample;
  int group, group2;
  int vec_ok;             // 16-byte path usable (alignment + numel % vec == 0)
};

struct KdParams {
  TapParam tap[PEA_KD_MAX_TAPS];
  int ntaps;
  int B;
  int total_chunks;
  float feature_weight;
  float grad_scale;
  int skip_nonfinite;
};

template <typename T> struct Vec16 { static constexpr int n = 16 / sizeof(T); };

template <typename T>
__device__ __forceinline__ void unpack16(const uint4& u, float* f) {
  if constexpr (sizeof(T) == 4) {
    f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y); f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
  } else {
    float2 a = unpack2<T>(u.x), b = unpack2<T>(u.y), c = unpack2<T>(u.z), d = unpack2<T>(u.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
  }
}
template <typename T>
__device__ __forceinline__ uint4 pack16(const float* f) {
  uint4 u;
  if constexpr (sizeof(T) == 4) {
    u.x = __float_as_uint(f[0]); u.y = __float_as_uint(f[1]); u.z = __float_as_uint(f[2]); u.w = __float_as_uint(f[3]);
  } else {
    u.x = pack2<T>(f[0], f[1]); u.y = pack2<T>(f[2], f[3]); u.z = pack2<T>(f[4], f[5]); u.w = pack2<T>(f[6], f[7]);
  }
  return u;
}

__device__ __forceinline__ float group_scale(int group, float feature_weight) { return group == 2 ? feature_weight : 1.0f; }

template <typename T>
__global__ void __launch_bounds__(kThreads) kd_mse_kernel(const __grid_constant__ KdParams p, float* __restrict__ partials) {
  // ---- which (tap, sample, chunk) is this CTA? -----------------------------------------------------------
  int ti = 0;
#pragma unroll 1
  for (int i = 1; i < p.ntaps; ++i)
    if ((int)blockIdx.x >= p.tap[i].chunk_begin) ti = i;
  const TapParam& tp = p.tap[ti];
  const int local = blockIdx.x - tp.chunk_begin;
  const int b = local / tp.chunks_per_sample;
  const int c = local - b * tp.chunks_per_sample;
  const long long base = (long long)b * tp.numel;
  const long long off0 = (long long)c * kChunk;
  const long long remaining = tp.numel - off0;
  const int n_here = remaining < kChunk ? (int)remaining : kChunk;

  const float w1 = tp.w ? tp.w[b] : 1.0f;
  const float w2 = (tp.t2 && tp.w2) ? tp.w2[b] : (tp.t2 ? 1.0f : 0.0f);
  const float inv = 1.0f / ((float)tp.numel * (float)p.B);
  const float gc1 = 2.0f * p.grad_scale * inv * w1 * group_scale(tp.group, p.feature_weight);
  const float gc2 = 2.0f * p.grad_scale * inv * w2 * group_scale(tp.group2, p.feature_weight);

  const T* __restrict__ s = static_cast<const T*>(tp.s) + base + off0;
  const T* __restrict__ t = static_cast<const T*>(tp.t) + base + off0;
  const T* __restrict__ t2 = tp.t2 ? static_cast<const T*>(tp.t2) + base + off0 : nullptr;
  T* __restrict__ g = tp.g ? static_cast<T*>(tp.g) + base + off0 : nullptr;

  float acc1 = 0.f, acc2 = 0.f;
  constexpr int V = Vec16<T>::n;
  if (tp.vec_ok) {
    constexpr int kIters = kElemsPerThread / V;  // 4 (16-bit) or 8 (fp32) 128-bit loads per tensor per thread
    uint4 sv[kIters], tv[kIters];
    // front-batch all loads (memory-level parallelism), then compute + store
#pragma unroll
    for (int i = 0; i < kIters; ++i) {
      const int e = (i * kThreads + threadIdx.x) * V;
      if (e < n_here) {
        sv[i] = ldg_stream(s + e);
        tv[i] = ldg_stream(t + e);
      }
    }
#pragma unroll
    for (int i = 0; i < kIters; ++i) {
      const int e = (i * kThreads + threadIdx.x) * V;
      if (e < n_here) {
        float fs[V], ft[V], fg[V];
        unpack16<T>(sv[i], fs);
        unpack16<T>(tv[i], ft);
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const float d = fs[j] - ft[j];
          acc1 = fmaf(d, d, acc1);
          fg[j] = gc1 * d;
        }
        if (t2) {
          float ft2[V];
          unpack16<T>(ldg_stream(t2 + e), ft2);
#pragma unroll
          for (int j = 0; j < V; ++j) {
            const float d = fs[j] - ft2[j];
            acc2 = fmaf(d, d, acc2);
            fg[j] = fmaf(gc2, d, fg[j]);
          }
        }
        if (g) stg_stream(g + e, pack16<T>(fg));
      }
    }
  } else {
    for (int e = threadIdx.x; e < n_here; e += kThreads) {
      const float fs = to_float(s[e]);
      const float d = fs - to_float(t[e]);
      acc1 = fmaf(d, d, acc1);
      float fg = gc1 * d;
      if (t2) {
        const float d2 = fs - to_float(t2[e]);
        acc2 = fmaf(d2, d2, acc2);
        fg = fmaf(gc2, d2, fg);
      }
      if (g) g[e] = from_float<T>(fg);
    }
  }

  // ---- block reduction, fixed order ------------------------------------------------------------------------
  acc1 = warp_sum(acc1);
  acc2 = warp_sum(acc2);
  __shared__ float red[2][kThreads / 32];
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = acc1;
    red[1][threadIdx.x >> 5] = acc2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, bsum = 0.f;
#pragma unroll
    for (int i = 0; i < kThreads / 32; ++i) { a += red[0][i]; bsum += red[1][i]; }
    partials[2 * blockIdx.x + 0] = a * w1 * inv;
    partials[2 * blockIdx.x + 1] = bsum * w2 * inv;
  }
}

// One CTA; sums the per-CTA partials of every tap in a fixed order and assembles the outputs.
__global__ void __launch_bounds__(kThreads) kd_finalize_kernel(const __grid_constant__ KdParams p,
                                                              const float* __restrict__ partials, float* __restrict__ out) {
  __shared__ float red[2][kThreads / 32];
  __shared__ float term[PEA_KD_MAX_TAPS][2];
  for (int ti = 0; ti < p.ntaps; ++ti) {
    const int begin = p.tap[ti].chunk_begin;
    const int end = begin + p.tap[ti].chunks_per_sample * p.B;
    float a = 0.f, b = 0.f;
    for (int i = begin + threadIdx.x; i < end; i += kThreads) {
      a += partials[2 * i];
      b += partials[2 * i + 1];
    }
    a = warp_sum(a);
    b = warp_sum(b);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float sa = 0.f, sb = 0.f;
      for (int i = 0; i < kThreads / 32; ++i) { sa += red[0][i]; sb += red[1][i]; }
      term[ti][0] = sa;
      term[ti][1] = sb;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    float L[3] = {0.f, 0.f, 0.f};
    for (int ti = 0; ti < p.ntaps; ++ti) {
      const TapParam& tp = p.tap[ti];
      const float v = term[ti][0] + term[ti][1];
      bool counted = true;
      // train_sd_zh.py:246,255,265: a feature term holding any NaN/Inf is left out of the sum
      if (p.skip_nonfinite && tp.group == 2 && !isfinite(v)) counted = false;
      if (counted) {
        L[tp.group] += term[ti][0];
        if (tp.t2) L[tp.group2] += term[ti][1];
      }
      out[4 + ti] = v;
      out[4 + p.ntaps + ti] = counted ? 1.0f : 0.0f;
    }
    out[0] = L[0];
    out[1] = L[1];
    out[2] = L[2];
    out[3] = L[0] + L[1] + p.feature_weight * L[2];
  }
}

// d_grad *= (upstream / prescale) * counted  (or = 0 for a skipped tap).  Exits at once when nothing has to change, which
// is the normal case: the fused pass already scaled by `prescale` (the trainer's loss scale) and autograd's upstream
// gradient IS that loss scale.
template <typename T>
__global__ void __launch_bounds__(kThreads) kd_scale_kernel(const __grid_constant__ KdParams p, const float* __restrict__ upstream,
                                                           const float* __restrict__ out) {
  const float up = (upstream ? upstream[0] : 1.0f) * p.grad_scale;   // grad_scale carries 1 / prescale here
  for (int ti = 0; ti < p.ntaps; ++ti) {
    const TapParam& tp = p.tap[ti];
    if (!tp.g) continue;
    const bool counted = out[4 + p.ntaps + ti] != 0.0f;
    if (counted && up == 1.0f) continue;
    T* g = static_cast<T*>(tp.g);
    const long long n = tp.numel * p.B;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads)
      g[i] = counted ? from_float<T>(to_float(g[i]) * up) : from_float<T>(0.0f);
  }
}

int fill_params(KdParams& p, const pea_kd_tap* taps, int ntaps, int B, pea_dtype dtype) {
  PEA_CHECK_ARG(taps && ntaps >= 1 && ntaps <= PEA_KD_MAX_TAPS, "kd: ntaps must be in [1,%d], got %d", PEA_KD_MAX_TAPS, ntaps);
  PEA_CHECK_ARG(B >= 1, "kd: B must be >= 1");
  const size_t es = dtype_size(dtype);
  const int vec = (int)(16 / es);
  long long total = 0;
  for (int i = 0; i < ntaps; ++i) {
    const pea_kd_tap& t = taps[i];
    PEA_CHECK_ARG(t.d_student && t.d_teacher, "kd: tap %d has a null student/teacher pointer", i);
    PEA_CHECK_ARG(t.numel_per_sample >= 1, "kd: tap %d numel_per_sample must be >= 1", i);
    PEA_CHECK_ARG(t.group >= 0 && t.group <= 2 && t.group2 >= 0 && t.group2 <= 2, "kd: tap %d group out of range", i);
    TapParam& q = p.tap[i];
    q.s = t.d_student; q.t = t.d_teacher; q.t2 = t.d_teacher2; q.g = t.d_grad; q.w = t.d_weight; q.w2 = t.d_weight2;
    q.numel = t.numel_per_sample;
    q.group = t.group; q.group2 = t.group2;
    q.chunks_per_sample = (int)((t.numel_per_sample + kChunk - 1) / kChunk);
    q.chunk_begin = (int)total;
    total += (long long)q.chunks_per_sample * B;
    auto al = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; };
    q.vec_ok = (t.numel_per_sample % vec == 0) && al(t.d_student) && al(t.d_teacher) && al(t.d_teacher2) && al(t.d_grad);
    PEA_CHECK_ARG(total < (1LL << 31), "kd: too many chunks");
  }
  p.ntaps = ntaps;
  p.B = B;
  p.total_chunks = (int)total;
  return PEA_OK;
}

}  // namespace
}  // namespace pea

using namespace pea;

extern "C" size_t pea_kd_workspace_bytes(const pea_kd_tap* taps, int ntaps, int B) {
  KdParams p{};
  if (fill_params(p, taps, ntaps, B, PEA_BF16) != PEA_OK) return 0;
  return (size_t)p.total_chunks * 2 * sizeof(float);
}

extern "C" int pea_kd_mse_fwd_bwd(const pea_kd_tap* taps, int ntaps, int B, pea_dtype dtype, float feature_weight,
                                  float grad_scale, int skip_nonfinite, void* d_workspace, float* d_out, void* stream) {
  KdParams p{};
  int rc = fill_params(p, taps, ntaps, B, dtype);
  if (rc != PEA_OK) return rc;
  PEA_CHECK_ARG(d_workspace && d_out, "kd: workspace/out must not be null");
  p.feature_weight = feature_weight;
  p.grad_scale = grad_scale;
  p.skip_nonfinite = skip_nonfinite;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* ws = static_cast<float*>(d_workspace);
  switch (dtype) {
    case PEA_F32: kd_mse_kernel<float><<<p.total_chunks, kThreads, 0, st>>>(p, ws); break;
    case PEA_F16: kd_mse_kernel<__half><<<p.total_chunks, kThreads, 0, st>>>(p, ws); break;
    case PEA_BF16: kd_mse_kernel<__nv_bfloat16><<<p.total_chunks, kThreads, 0, st>>>(p, ws); break;
    default: PEA_CHECK_ARG(false, "kd: bad dtype %d", (int)dtype);
  }
  PEA_LAUNCH_CHECK();
  kd_finalize_kernel<<<1, kThreads, 0, st>>>(p, ws, d_out);
  PEA_LAUNCH_CHECK();
  return PEA_OK;
}

extern "C" int pea_kd_scale_grads(const pea_kd_tap* taps, int ntaps, int B, pea_dtype dtype, const float* d_upstream,
                                  float inv_prescale, const float* d_out, void* stream) {
  KdParams p{};
  int rc = fill_params(p, taps, ntaps, B, dtype);
  if (rc != PEA_OK) return rc;
  PEA_CHECK_ARG(d_out, "kd: out must not be null");
  PEA_CHECK_ARG(inv_prescale > 0.0f, "kd: inv_prescale must be > 0");
  p.grad_scale = inv_prescale;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = num_sms() * 8;
  switch (dtype) {
    case PEA_F32: kd_scale_kernel<float><<<grid, kThreads, 0, st>>>(p, d_upstream, d_out); break;
    case PEA_F16: kd_scale_kernel<__half><<<grid, kThreads, 0, st>>>(p, d_upstream, d_out); break;
    case PEA_BF16: kd_scale_kernel<__nv_bfloat16><<<grid, kThreads, 0, st>>>(p, d_upstream, d_out); break;
    default: PEA_CHECK_ARG(false, "kd: bad dtype %d", (int)dtype);
  }
  PEA_LAUNCH_CHECK();
  return PEA_OK;
}
