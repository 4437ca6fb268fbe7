// rowops.cu -- the adapter's row-wise pieces and the CFG-dropout select.
//   LayerNorm(in_dim)                      train_sdxl_zh.py:47,60 / train_sd_zh.py:44,54
//   torch.mean(x, 1)                       train_sdxl_zh.py:66
//   where(prompt_mask, uncond, cond)       train_sdxl_zh.py:392-395
// plus operand splitting (fp32 -> bf16 planes) for the fp32-accurate GEMM mode and the bias/affine gradients.
// All of these touch <= a few MB per step (M = 2*B*77 rows) -- except the plane split of the all-layer dK|dV accumulator
// (410 MB).  Round 2: the ncu launch list of a step (profiles/r2_launches_step_graph.txt) showed them costing a third of the
// adapter's forward+backward (LayerNorm backward 38 us on 32 CTAs, column sum 20 us, LayerNorm forward 20 us, 7-12 us per
// weight cast with a 64-bit division per element), so the hot ones got vectorised / row-parallel forms.  The affine and
// bias gradients are now reduced with fp32 atomics over row chunks (order-dependent in the last bits, like dK / dV).
#include "common.cuh"

namespace pea {
namespace {

template <typename T> __device__ __forceinline__ float ld_as_float(const void* p, long long i) { return to_float(static_cast<const T*>(p)[i]); }

__device__ __forceinline__ float load_any(const void* p, pea_dtype dt, long long i) {
  switch (dt) {
    case PEA_F32: return static_cast<const float*>(p)[i];
    case PEA_F16: return __half2float(static_cast<const __half*>(p)[i]);
    default: return __bfloat162float(static_cast<const __nv_bfloat16*>(p)[i]);
  }
}
__device__ __forceinline__ void store_any(void* p, pea_dtype dt, long long i, float v) {
  switch (dt) {
    case PEA_F32: static_cast<float*>(p)[i] = v; break;
    case PEA_F16: static_cast<__half*>(p)[i] = __float2half_rn(v); break;
    default: static_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v); break;
  }
}

struct Planes { void* p[3]; };

// v -> S planes whose sum reproduces v to ~8*S mantissa bits; plane s holds the rounding residual of planes < s.
__device__ __forceinline__ void store_split(const Planes& pl, int nsplit, pea_dtype dt, long long i, float v) {
  float r = v;
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    if (s < nsplit) {
      float q;
      if (dt == PEA_F16) { __half h = __float2half_rn(r); static_cast<__half*>(pl.p[s])[i] = h; q = __half2float(h); }
      else { __nv_bfloat16 h = __float2bfloat16_rn(r); static_cast<__nv_bfloat16*>(pl.p[s])[i] = h; q = __bfloat162float(h); }
      r -= q;
    }
  }
}

// one warp per row
__global__ void layernorm_fwd_kernel(const void* __restrict__ x, pea_dtype xdt, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, int M, int D, float eps, Planes y, int nsplit,
                                     pea_dtype pdt, float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  pdl_wait();
  pdl_launch_dependents();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const long long base = (long long)row * D;
  float s = 0.f;
  for (int i = lane; i < D; i += 32) s += load_any(x, xdt, base + i);
  const float mean = warp_sum(s) / (float)D;
  float v = 0.f;
  for (int i = lane; i < D; i += 32) { const float d = load_any(x, xdt, base + i) - mean; v = fmaf(d, d, v); }
  const float rstd = rsqrtf(warp_sum(v) / (float)D + eps);
  for (int i = lane; i < D; i += 32) {
    const float xh = (load_any(x, xdt, base + i) - mean) * rstd;
    store_split(y, nsplit, pdt, base + i, fmaf(xh, gamma[i], beta[i]));
  }
  if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
}

// Same, fp32 rows with D = 128 * NV (<= 2048): the row is read ONCE with 128-bit loads and stays in registers for the mean,
// the variance and the normalised output (the general kernel above makes three latency-bound passes of scalar loads).
template <int NV>
__global__ void layernorm_fwd_vec_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                         int M, float eps, Planes y, int nsplit, pea_dtype pdt, float* __restrict__ mean_out,
                                         float* __restrict__ rstd_out) {
  pdl_wait();
  pdl_launch_dependents();
  constexpr int D = NV * 128;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (long long)row * D);
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) { v[i] = xr[i * 32 + lane]; s += (v[i].x + v[i].y) + (v[i].z + v[i].w); }
  const float mean = warp_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c0 = (i * 32 + lane) * 4;
    const float4 g = *reinterpret_cast<const float4*>(gamma + c0), b = *reinterpret_cast<const float4*>(beta + c0);
    const long long base = (long long)row * D + c0;
    store_split(y, nsplit, pdt, base + 0, fmaf((v[i].x - mean) * rstd, g.x, b.x));
    store_split(y, nsplit, pdt, base + 1, fmaf((v[i].y - mean) * rstd, g.y, b.y));
    store_split(y, nsplit, pdt, base + 2, fmaf((v[i].z - mean) * rstd, g.z, b.z));
    store_split(y, nsplit, pdt, base + 3, fmaf((v[i].w - mean) * rstd, g.w, b.w));
  }
  if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
}

// dgamma[d] += sum_m dy*xhat ; dbeta[d] += sum_m dy.  block (32, 8): 32 columns; grid.y chunks of rows, each chunk's partial
// sums added with one fp32 atomic per column (the caller zero-fills or really accumulates).
__global__ void layernorm_bwd_affine_kernel(const void* __restrict__ x, pea_dtype xdt, const float* __restrict__ dy,
                                            const float* __restrict__ mean, const float* __restrict__ rstd, int M, int D,
                                            float* __restrict__ dgamma, float* __restrict__ dbeta) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float sg[8][33], sb[8][33];
  const int col = blockIdx.x * 32 + threadIdx.x;
  const int rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int m_begin = blockIdx.y * rows_per, m_end = min(M, m_begin + rows_per);
  float ag = 0.f, ab = 0.f;
  if (col < D) {
    for (int m = m_begin + threadIdx.y; m < m_end; m += 8) {
      const float g = dy[(long long)m * D + col];
      const float xh = (load_any(x, xdt, (long long)m * D + col) - mean[m]) * rstd[m];
      ag = fmaf(g, xh, ag);
      ab += g;
    }
  }
  sg[threadIdx.y][threadIdx.x] = ag;
  sb[threadIdx.y][threadIdx.x] = ab;
  __syncthreads();
  if (threadIdx.y == 0 && col < D) {
    float tg = 0.f, tb = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { tg += sg[i][threadIdx.x]; tb += sb[i][threadIdx.x]; }
    atomicAdd(dgamma + col, tg);
    atomicAdd(dbeta + col, tb);
  }
}

// dx = rstd * (dy*gamma - mean(dy*gamma) - xhat * mean(dy*gamma*xhat)); one warp per row
__global__ void layernorm_bwd_dx_kernel(const void* __restrict__ x, pea_dtype xdt, const float* __restrict__ dy,
                                        const float* __restrict__ gamma, const float* __restrict__ mean,
                                        const float* __restrict__ rstd, int M, int D, float* __restrict__ dx) {
  pdl_wait();
  pdl_launch_dependents();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const long long base = (long long)row * D;
  const float mu = mean[row], rs = rstd[row];
  float c1 = 0.f, c2 = 0.f;
  for (int i = lane; i < D; i += 32) {
    const float g = dy[base + i] * gamma[i];
    const float xh = (load_any(x, xdt, base + i) - mu) * rs;
    c1 += g;
    c2 = fmaf(g, xh, c2);
  }
  c1 = warp_sum(c1) / (float)D;
  c2 = warp_sum(c2) / (float)D;
  for (int i = lane; i < D; i += 32) {
    const float g = dy[base + i] * gamma[i];
    const float xh = (load_any(x, xdt, base + i) - mu) * rs;
    dx[base + i] = rs * (g - c1 - xh * c2);
  }
}

__global__ void split_planes_kernel(const void* __restrict__ x, pea_dtype xdt, long long rows, long long cols, long long ld,
                                    Planes pl, int nsplit, pea_dtype pdt) {
  pdl_wait();
  pdl_launch_dependents();
  const long long n = rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols, c = i - r * cols;
    store_split(pl, nsplit, pdt, i, load_any(x, xdt, r * ld + c));
  }
}

// Contiguous fp32 source (ld == cols, 32-byte aligned, n % 8 == 0): 8 elements per thread per step, 128-bit loads, one
// 128-bit store per plane, no index arithmetic beyond the grid stride.
struct SplitItem { const float* x; long long n8; Planes pl; };
constexpr int kSplitBatch = 8;
struct SplitBatch { SplitItem it[kSplitBatch]; };     // blockIdx.y selects the tensor (pea_split_planes_multi: one launch)

__global__ void split_planes_vec_kernel(const __grid_constant__ SplitBatch batch, int nsplit, pea_dtype pdt) {
  pdl_wait();
  pdl_launch_dependents();
  const float* __restrict__ x = batch.it[blockIdx.y].x;
  const long long n8 = batch.it[blockIdx.y].n8;
  const Planes pl = batch.it[blockIdx.y].pl;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(x)[2 * i], b = reinterpret_cast<const float4*>(x)[2 * i + 1];
    float r[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      if (s < nsplit) {
        uint4 u;
        if (pdt == PEA_F16) {
          u = make_uint4(pack2<__half>(r[0], r[1]), pack2<__half>(r[2], r[3]), pack2<__half>(r[4], r[5]), pack2<__half>(r[6], r[7]));
          if (s + 1 < nsplit) {
            const float2 q0 = unpack2<__half>(u.x), q1 = unpack2<__half>(u.y), q2 = unpack2<__half>(u.z), q3 = unpack2<__half>(u.w);
            r[0] -= q0.x; r[1] -= q0.y; r[2] -= q1.x; r[3] -= q1.y; r[4] -= q2.x; r[5] -= q2.y; r[6] -= q3.x; r[7] -= q3.y;
          }
        } else {
          u = make_uint4(pack2<__nv_bfloat16>(r[0], r[1]), pack2<__nv_bfloat16>(r[2], r[3]), pack2<__nv_bfloat16>(r[4], r[5]),
                         pack2<__nv_bfloat16>(r[6], r[7]));
          if (s + 1 < nsplit) {
            const float2 q0 = unpack2<__nv_bfloat16>(u.x), q1 = unpack2<__nv_bfloat16>(u.y), q2 = unpack2<__nv_bfloat16>(u.z),
                         q3 = unpack2<__nv_bfloat16>(u.w);
            r[0] -= q0.x; r[1] -= q0.y; r[2] -= q1.x; r[3] -= q1.y; r[4] -= q2.x; r[5] -= q2.y; r[6] -= q3.x; r[7] -= q3.y;
          }
        }
        reinterpret_cast<uint4*>(pl.p[s])[i] = u;
      }
    }
  }
}

__global__ void mean_tokens_kernel(const float* __restrict__ x, const void* __restrict__ res, pea_dtype rdt, int L, int D,
                                   float* __restrict__ out) {
  pdl_wait();
  pdl_launch_dependents();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (n >= D) return;
  float s4[4] = {0.f, 0.f, 0.f, 0.f};      // four loads in flight (one serial chain of 77 dependent adds took 9 us)
  const long long base = (long long)b * L * D + n;
  int l = 0;
  for (; l + 4 <= L; l += 4) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long i = base + (long long)(l + j) * D;
      float v = x[i];
      if (res) v += load_any(res, rdt, i);
      s4[j] += v;
    }
  }
  for (; l < L; ++l) {
    const long long i = base + (long long)l * D;
    float v = x[i];
    if (res) v += load_any(res, rdt, i);
    s4[0] += v;
  }
  out[(long long)b * D + n] = ((s4[0] + s4[1]) + (s4[2] + s4[3])) / (float)L;
}

__global__ void colsum_accum_kernel(const void* __restrict__ x, pea_dtype xdt, int M, int N, long long ld, float* __restrict__ out) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float sm[8][33];
  const int col = blockIdx.x * 32 + threadIdx.x;
  const int rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int m_begin = blockIdx.y * rows_per, m_end = min(M, m_begin + rows_per);
  float a = 0.f;
  if (col < N)
    for (int m = m_begin + threadIdx.y; m < m_end; m += 8) a += load_any(x, xdt, (long long)m * ld + col);
  sm[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && col < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sm[i][threadIdx.x];
    atomicAdd(out + col, t);
  }
}

template <typename T>
__global__ void cfg_select_kernel(const unsigned char* __restrict__ mask, const T* __restrict__ u, const T* __restrict__ c,
                                  T* __restrict__ out, long long row_elems) {
  const int b = blockIdx.y;
  const T* src = mask[b] ? u : c;
  const long long base = (long long)b * row_elems;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < row_elems; i += (long long)gridDim.x * blockDim.x)
    out[base + i] = src[base + i];
}

__global__ void add_noise_kernel(const void* __restrict__ lat, pea_dtype ldt, const float* __restrict__ raw, const float* __restrict__ offs,
                                 float noise_offset, const float* __restrict__ acp, int T, const long long* __restrict__ t, int C,
                                 long long HW, void* __restrict__ noise, void* __restrict__ noisy, pea_dtype odt) {
  const int b = blockIdx.z, c = blockIdx.y;
  long long ts = t[b];
  ts = ts < 0 ? 0 : (ts >= T ? T - 1 : ts);
  const float a = acp[ts];
  const float sa = sqrtf(a), sb = sqrtf(1.0f - a);
  const float off = offs ? noise_offset * offs[(long long)b * C + c] : 0.f;
  const long long base = ((long long)b * C + c) * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x) {
    // the reference rounds the noise to the UNet dtype BEFORE add_noise (train_sdxl_zh.py:317): do the same
    float n = raw[base + i] + off;
    store_any(noise, odt, base + i, n);
    n = load_any(noise, odt, base + i);
    store_any(noisy, odt, base + i, sa * load_any(lat, ldt, base + i) + sb * n);
  }
}

}  // namespace
}  // namespace pea

using namespace pea;

extern "C" int pea_add_noise(const void* d_latents, pea_dtype lat_dtype, const float* d_noise_raw, const float* d_offs, float noise_offset,
                             const float* d_acp, int T, const long long* d_t, int B, int C, long long HW, void* d_noise, void* d_noisy,
                             pea_dtype out_dtype, void* stream) {
  PEA_CHECK_ARG(d_latents && d_noise_raw && d_acp && d_t && d_noise && d_noisy, "add_noise: null pointer");
  PEA_CHECK_ARG(B >= 1 && C >= 1 && HW >= 1 && T >= 1 && C <= 65535 && B <= 65535, "add_noise: bad shape");
  const int gx = (int)((HW + 255) / 256 < 64 ? (HW + 255) / 256 : 64);
  add_noise_kernel<<<dim3(gx, C, B), 256, 0, (cudaStream_t)stream>>>(d_latents, lat_dtype, d_noise_raw, d_offs, noise_offset, d_acp, T,
                                                                    d_t, C, HW, d_noise, d_noisy, out_dtype);
  PEA_LAUNCH_CHECK();
  return PEA_OK;
}

extern "C" int pea_layernorm_fwd(const void* d_x, pea_dtype x_dtype, const float* d_gamma, const float* d_beta, int M, int D,
                                 float eps, void* const d_y_planes[3], int nsplit, pea_dtype plane_dtype, float* d_mean,
                                 float* d_rstd, unsigned flags, void* stream) {
  PEA_CHECK_ARG(d_x && d_gamma && d_beta && d_y_planes && d_mean && d_rstd, "layernorm_fwd: null pointer");
  PEA_CHECK_ARG(M >= 1 && D >= 1 && nsplit >= 1 && nsplit <= 3, "layernorm_fwd: bad M/D/nsplit");
  PEA_CHECK_ARG(plane_dtype != PEA_F32, "layernorm_fwd: planes are 16-bit");
  Planes pl{};
  for (int s = 0; s < nsplit; ++s) { PEA_CHECK_ARG(d_y_planes[s], "layernorm_fwd: null plane %d", s); pl.p[s] = d_y_planes[s]; }
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const bool vec = x_dtype == PEA_F32 && D % 128 == 0 && D <= 2048 && al16(d_x) && al16(d_gamma) && al16(d_beta);
  cudaStream_t st = (cudaStream_t)stream;
  const float* xf = static_cast<const float*>(d_x);
  const bool pdl = (flags & PEA_LAUNCH_PDL) != 0;
#define PEA_LN_VEC(NV) PEA_LAUNCH_EX(layernorm_fwd_vec_kernel<NV>, (M + 3) / 4, 128, 0, st, pdl, xf, d_gamma, d_beta, M, eps, pl, nsplit, plane_dtype, d_mean, d_rstd)
  if (vec && D == 1024) PEA_LN_VEC(8);
  else if (vec && D == 768) PEA_LN_VEC(6);
  else if (vec && D == 2048) PEA_LN_VEC(16);
  else if (vec && D == 128) PEA_LN_VEC(1);
  else PEA_LAUNCH_EX(layernorm_fwd_kernel, (M + 3) / 4, 128, 0, st, pdl, d_x, x_dtype, d_gamma, d_beta, M, D, eps, pl, nsplit, plane_dtype, d_mean, d_rstd);
#undef PEA_LN_VEC
  PEA_LAUNCH_CHECK();
  return PEA_OK;
}

extern "C" int pea_layernorm_bwd(const void* d_x, pea_dtype x_dtype, const float* d_dy, const float* d_gamma, const float* d_mean,
                                 const float* d_rstd, int M, int D, float* d_dgamma, float* d_dbeta, float* d_dx, unsigned flags,
                                 void* stream) {
  PEA_CHECK_ARG(d_x && d_dy && d_gamma && d_mean && d_rstd && d_dgamma && d_dbeta, "layernorm_bwd: null pointer");
  PEA_CHECK_ARG(M >= 1 && D >= 1, "layernorm_bwd: bad M/D");
  // enough row chunks to cover the machine (D = 1024: 32 column blocks x 8 row chunks); at least 64 rows per chunk
  const int cb = (D + 31) / 32;
  int rc = (2 * num_sms() + cb - 1) / cb;
  if (rc > (M + 63) / 64) rc = (M + 63) / 64;
  if (rc < 1) rc = 1;
  const bool pdl = (flags & PEA_LAUNCH_PDL) != 0;
  PEA_LAUNCH_EX(layernorm_bwd_affine_kernel, dim3(cb, rc), dim3(32, 8), 0, (cudaStream_t)stream, pdl, d_x, x_dtype, d_dy, d_mean,
                        d_rstd, M, D, d_dgamma, d_dbeta);
  if (d_dx)
    PEA_LAUNCH_EX(layernorm_bwd_dx_kernel, (M + 3) / 4, 128, 0, (cudaStream_t)stream, pdl, d_x, x_dtype, d_dy, d_gamma, d_mean,
                          d_rstd, M, D, d_dx);
  return PEA_OK;
}

extern "C" int pea_split_planes(const void* d_x, pea_dtype x_dtype, long long rows, long long cols, long long ld,
                                void* const d_planes[3], int nsplit, pea_dtype plane_dtype, unsigned flags, void* stream) {
  PEA_CHECK_ARG(d_x && d_planes && rows >= 1 && cols >= 1 && ld >= cols, "split_planes: bad arguments");
  PEA_CHECK_ARG(nsplit >= 1 && nsplit <= 3 && plane_dtype != PEA_F32, "split_planes: bad nsplit/plane dtype");
  Planes pl{};
  for (int s = 0; s < nsplit; ++s) { PEA_CHECK_ARG(d_planes[s], "split_planes: null plane %d", s); pl.p[s] = d_planes[s]; }
  const long long n = rows * cols;
  const bool pdl = (flags & PEA_LAUNCH_PDL) != 0;
  bool vec = x_dtype == PEA_F32 && ld == cols && n % 8 == 0 && (reinterpret_cast<uintptr_t>(d_x) & 31) == 0;
  for (int s = 0; s < nsplit; ++s) vec = vec && (reinterpret_cast<uintptr_t>(d_planes[s]) & 15) == 0;
  if (vec) {
    const long long n8 = n / 8;
    const int grid = (int)((n8 + 255) / 256 < (long long)num_sms() * 16 ? (n8 + 255) / 256 : (long long)num_sms() * 16);
    SplitBatch batch{};
    batch.it[0] = SplitItem{static_cast<const float*>(d_x), n8, pl};
    PEA_LAUNCH_EX(split_planes_vec_kernel, grid, 256, 0, (cudaStream_t)stream, pdl, batch, nsplit, plane_dtype);
    return PEA_OK;
  }
  const int grid = (int)((n + 255) / 256 < (long long)num_sms() * 16 ? (n + 255) / 256 : (long long)num_sms() * 16);
  PEA_LAUNCH_EX(split_planes_kernel, grid, 256, 0, (cudaStream_t)stream, pdl, d_x, x_dtype, rows, cols, ld, pl, nsplit, plane_dtype);
  return PEA_OK;
}

extern "C" int pea_split_planes_multi(const pea_split_item* items, int count, pea_dtype x_dtype, int nsplit, pea_dtype plane_dtype,
                                      unsigned flags, void* stream) {
  PEA_CHECK_ARG(items && count >= 1, "split_planes_multi: bad arguments");
  PEA_CHECK_ARG(nsplit >= 1 && nsplit <= 3 && plane_dtype != PEA_F32, "split_planes_multi: bad nsplit/plane dtype");
  // one launch when every tensor takes the vector path (dense fp32, 8 | n, aligned) -- the adapter's weights; else one call each
  bool vec = x_dtype == PEA_F32 && count <= kSplitBatch;
  long long n8_max = 0;
  SplitBatch batch{};
  for (int i = 0; i < count && vec; ++i) {
    const pea_split_item& it = items[i];
    PEA_CHECK_ARG(it.d_x && it.rows >= 1 && it.cols >= 1 && it.ld >= it.cols, "split_planes_multi: bad item %d", i);
    const long long n = it.rows * it.cols;
    vec = it.ld == it.cols && n % 8 == 0 && (reinterpret_cast<uintptr_t>(it.d_x) & 31) == 0;
    for (int s = 0; s < nsplit; ++s) {
      PEA_CHECK_ARG(it.d_planes[s], "split_planes_multi: item %d: null plane %d", i, s);
      vec = vec && (reinterpret_cast<uintptr_t>(it.d_planes[s]) & 15) == 0;
      batch.it[i].pl.p[s] = it.d_planes[s];
    }
    batch.it[i].x = static_cast<const float*>(it.d_x);
    batch.it[i].n8 = n / 8;
    if (n / 8 > n8_max) n8_max = n / 8;
  }
  if (!vec) {
    for (int i = 0; i < count; ++i) {
      const int rc = pea_split_planes(items[i].d_x, x_dtype, items[i].rows, items[i].cols, items[i].ld, items[i].d_planes, nsplit,
                                      plane_dtype, flags, stream);
      if (rc != PEA_OK) return rc;
    }
    return PEA_OK;
  }
  long long gx = (n8_max + 255) / 256;
  const long long cap = ((long long)num_sms() * 16 + count - 1) / count;
  if (gx > cap) gx = cap;
  PEA_LAUNCH_EX(split_planes_vec_kernel, dim3((unsigned)gx, (unsigned)count), dim3(256), 0, (cudaStream_t)stream,
                     (flags & PEA_LAUNCH_PDL) != 0, batch, nsplit, plane_dtype);
  return PEA_OK;
}

extern "C" int pea_mean_tokens(const float* d_x, const void* d_res, pea_dtype res_dtype, int B, int L, int D, float* d_out,
                               unsigned flags, void* stream) {
  PEA_CHECK_ARG(d_x && d_out && B >= 1 && L >= 1 && D >= 1, "mean_tokens: bad arguments");
  PEA_LAUNCH_EX(mean_tokens_kernel, dim3((D + 127) / 128, B), dim3(128), 0, (cudaStream_t)stream, (flags & PEA_LAUNCH_PDL) != 0,
                        d_x, d_res, res_dtype, L, D, d_out);
  return PEA_OK;
}

extern "C" int pea_colsum_accum(const void* d_x, pea_dtype x_dtype, int M, int N, long long ld, float* d_out, unsigned flags,
                                void* stream) {
  PEA_CHECK_ARG(d_x && d_out && M >= 1 && N >= 1 && ld >= N, "colsum_accum: bad arguments");
  const int cb = (N + 31) / 32;
  int rc = (2 * num_sms() + cb - 1) / cb;
  if (rc > (M + 63) / 64) rc = (M + 63) / 64;
  if (rc < 1) rc = 1;
  PEA_LAUNCH_EX(colsum_accum_kernel, dim3(cb, rc), dim3(32, 8), 0, (cudaStream_t)stream, (flags & PEA_LAUNCH_PDL) != 0, d_x, x_dtype,
                        M, N, ld, d_out);
  return PEA_OK;
}

extern "C" int pea_cfg_select(const unsigned char* d_mask, const void* d_uncond, const void* d_cond, void* d_out, int B,
                              long long row_elems, pea_dtype dtype, void* stream) {
  PEA_CHECK_ARG(d_mask && d_uncond && d_cond && d_out && B >= 1 && row_elems >= 1, "cfg_select: bad arguments");
  const int gx = (int)((row_elems + 255) / 256 < 64 ? (row_elems + 255) / 256 : 64);
  if (dtype == PEA_F32)
    cfg_select_kernel<float><<<dim3(gx, B), 256, 0, (cudaStream_t)stream>>>(d_mask, (const float*)d_uncond, (const float*)d_cond,
                                                                           (float*)d_out, row_elems);
  else
    cfg_select_kernel<uint16_t><<<dim3(gx, B), 256, 0, (cudaStream_t)stream>>>(d_mask, (const uint16_t*)d_uncond,
                                                                              (const uint16_t*)d_cond, (uint16_t*)d_out, row_elems);
  PEA_LAUNCH_CHECK();
  return PEA_OK;
}
