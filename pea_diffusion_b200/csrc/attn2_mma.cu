// attn2_mma.cu -- fused cross-attention core (kernel K3), forward and backward.
//
// Replaces F.scaled_dot_product_attention inside diffusers' AttnProcessor2_0 as reached from
// train_sdxl_zh.py:397 (student, autograd), :415 (teacher, no_grad) and tests/test_sdxl_zh.py:384-391.
//
// Shape regime: <= 80 keys (77 CLIP tokens), head_dim 64 (SDXL) or 40/80/160 (SD1.5), 1k-16k queries.  All keys
// of one (batch, head) fit one tile, so
//   * forward  : K and V (2 x 80 x d) stay resident in shared memory; Q streams through once, O streams out once;
//                softmax is single-pass (no online rescaling), row max / row sum by quad shuffles.
//   * backward : S and P are recomputed from Q, K; delta = rowsum(P o dP) is local to the row, so neither O nor a
//                log-sum-exp is re-read: traffic is read Q, read dO, write dQ (+ tiny K, V, dK, dV).
// Tensor-core math uses warp-level mma.sync.m16n8k16 (bf16/fp16 in, fp32 accumulate); see DESIGN.md for why this
// kernel (HBM-bound, 80-wide key tile) is the first implementation and where the tcgen05 variant stands.
//
// Algorithmic bytes (e = 2): fwd 2*B*N*C*e + 2*B*L*C*e ; bwd 3*B*N*C*e + 2*B*L*C*e + 2*B*L*C*4.
#include <stdlib.h>

#include "common.cuh"

namespace pea {
namespace {

constexpr int kLP = 80;          // padded key count (10 n8 tiles / 5 k16 steps)
constexpr int kNT = kLP / 8;     // n-tiles of S
constexpr int kPK = kLP / 16;    // k-steps of P*V
constexpr int kLDP = kLP + 8;    // row stride (elements) of the P / dS staging tiles
constexpr int kWarps = 4;

struct AttnParams {
  int B, N, L, H, d;
  float scale_log2;  // scale * log2(e)
  float scale;
  const void* q; long long ld_q;
  const void* k; const void* v; long long ld_kv, bs_kv;
  void* o; long long ld_o;
  const void* dout; long long ld_do;
  void* dq; long long ld_dq;
  float* dk; float* dv; long long ld_dkv, bs_dkv;
  int rows_per_cta;  // backward: query rows owned by one CTA (multiple of 64)
};

// ---- PTX wrappers ------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
template <typename T> __device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <> __device__ __forceinline__ void mma16816<__nv_bfloat16>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <> __device__ __forceinline__ void mma16816<__half>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Stage a [rows_valid x d] slab of a row-major global tensor into a padded smem tile [rows][LDS]; rows/cols beyond
// the valid range are zero-filled.  16-byte cp.async, coalesced along the row.
template <typename T, int D, int LDS>
__device__ __forceinline__ void stage_tile(T* smem, const T* gptr, long long ld, int rows, int rows_valid, int d, int tid, int nthreads) {
  constexpr int CH = D / 8;
  for (int idx = tid; idx < rows * CH; idx += nthreads) {
    const int r = idx / CH, c = idx - r * CH;
    T* dst = smem + r * LDS + c * 8;
    if (r < rows_valid && c * 8 < d) cp_async16(smem_u32(dst), gptr + (long long)r * ld + c * 8);
    else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
  }
}

// S[16 x 80] (+)= A[16 x D] * Bt[80 x D]^T for one warp; A rows start at a_row0 of As, Bt is Ks or Vs.
template <typename T, int D, int LDS>
__device__ __forceinline__ void warp_qk(float (&s)[kNT][4], const T* As, int a_row0, const T* Bs, int lane) {
#pragma unroll
  for (int ks = 0; ks < D / 16; ++ks) {
    uint32_t a[4];
    ldsm_x4(smem_u32(As + (a_row0 + (lane & 15)) * LDS + ks * 16 + (lane >> 4) * 8), a[0], a[1], a[2], a[3]);
#pragma unroll
    for (int n2 = 0; n2 < kNT / 2; ++n2) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4(smem_u32(Bs + (n2 * 16 + (lane & 7) + ((lane >> 4) << 3)) * LDS + ks * 16 + ((lane >> 3) & 1) * 8), b0, b1, b2, b3);
      mma16816<T>(s[2 * n2], a, b0, b1);
      mma16816<T>(s[2 * n2 + 1], a, b2, b3);
    }
  }
}

// O[16 x D] += P[16 x 80] * Bs[80 x D] for one warp, P given as packed A fragments (5 k-steps).
template <typename T, int D, int LDS>
__device__ __forceinline__ void warp_pv(float (&o)[D / 8][4], const uint32_t (&pa)[kPK][4], const T* Bs, int lane) {
#pragma unroll
  for (int kt = 0; kt < kPK; ++kt) {
#pragma unroll
    for (int d2 = 0; d2 < D / 16; ++d2) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(smem_u32(Bs + (kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + d2 * 16 + (lane >> 4) * 8), b0, b1, b2, b3);
      mma16816<T>(o[2 * d2], pa[kt], b0, b1);
      mma16816<T>(o[2 * d2 + 1], pa[kt], b2, b3);
    }
  }
}

// Row softmax of a warp's S fragment (rows lane/4 and lane/4+8), keys >= L masked.  Leaves normalised P in s.
__device__ __forceinline__ void warp_softmax(float (&s)[kNT][4], int L, float scale_log2, int lane) {
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < kNT; ++nt) {
    const int c = nt * 8 + 2 * (lane & 3);
    if (c >= L) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
    if (c + 1 >= L) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
    m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
    m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
  }
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
  const float o0 = m0 * scale_log2, o1 = m1 * scale_log2;
  float l0 = 0.f, l1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < kNT; ++nt) {
    s[nt][0] = exp2f(fmaf(s[nt][0], scale_log2, -o0)); s[nt][1] = exp2f(fmaf(s[nt][1], scale_log2, -o0));
    s[nt][2] = exp2f(fmaf(s[nt][2], scale_log2, -o1)); s[nt][3] = exp2f(fmaf(s[nt][3], scale_log2, -o1));
    l0 += s[nt][0] + s[nt][1];
    l1 += s[nt][2] + s[nt][3];
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
#pragma unroll
  for (int nt = 0; nt < kNT; ++nt) { s[nt][0] *= i0; s[nt][1] *= i0; s[nt][2] *= i1; s[nt][3] *= i1; }
}

template <typename T>
__device__ __forceinline__ void pack_afrag(uint32_t (&pa)[kPK][4], const float (&s)[kNT][4]) {
#pragma unroll
  for (int kt = 0; kt < kPK; ++kt) {
    pa[kt][0] = pack2<T>(s[2 * kt][0], s[2 * kt][1]);
    pa[kt][1] = pack2<T>(s[2 * kt][2], s[2 * kt][3]);
    pa[kt][2] = pack2<T>(s[2 * kt + 1][0], s[2 * kt + 1][1]);
    pa[kt][3] = pack2<T>(s[2 * kt + 1][2], s[2 * kt + 1][3]);
  }
}

// Warp-private 16 x D accumulator -> smem rows (as T) -> coalesced 16-byte global stores.
template <typename T, int D, int LDS>
__device__ __forceinline__ void warp_store_rows(const float (&acc)[D / 8][4], T* stage, int row0, T* gptr, long long ld,
                                                int rows_valid, int d, int lane) {
#pragma unroll
  for (int dt = 0; dt < D / 8; ++dt) {
    const int c = dt * 8 + 2 * (lane & 3);
    *reinterpret_cast<uint32_t*>(stage + (row0 + (lane >> 2)) * LDS + c) = pack2<T>(acc[dt][0], acc[dt][1]);
    *reinterpret_cast<uint32_t*>(stage + (row0 + (lane >> 2) + 8) * LDS + c) = pack2<T>(acc[dt][2], acc[dt][3]);
  }
  __syncwarp();
  constexpr int CH = D / 8;
  for (int idx = lane; idx < 16 * CH; idx += 32) {
    const int r = idx / CH, c = idx - r * CH;
    if (row0 + r < rows_valid && c * 8 < d)
      stg_stream(gptr + (long long)(row0 + r) * ld + c * 8, *reinterpret_cast<const uint4*>(stage + (row0 + r) * LDS + c * 8));
  }
}

// =============================================================================================================
// forward: grid (ceil(N / 64), H, B), 4 warps, one m16 tile per warp
// =============================================================================================================
template <typename T, int D>
__global__ void __launch_bounds__(kWarps * 32) attn2_fwd_kernel(const __grid_constant__ AttnParams p) {
  constexpr int LDS = D + 8;
  constexpr int ROWS = kWarps * 16;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* Ks = reinterpret_cast<T*>(smem_raw);
  T* Vs = Ks + kLP * LDS;
  T* Qs = Vs + kLP * LDS;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = blockIdx.x * ROWS;
  const int rows_valid = min(ROWS, p.N - q0);

  const T* kg = static_cast<const T*>(p.k) + (long long)b * p.bs_kv + (long long)h * p.d;
  const T* vg = static_cast<const T*>(p.v) + (long long)b * p.bs_kv + (long long)h * p.d;
  const T* qg = static_cast<const T*>(p.q) + ((long long)b * p.N + q0) * p.ld_q + (long long)h * p.d;
  stage_tile<T, D, LDS>(Qs, qg, p.ld_q, ROWS, rows_valid, p.d, tid, kWarps * 32);
  stage_tile<T, D, LDS>(Ks, kg, p.ld_kv, kLP, p.L, p.d, tid, kWarps * 32);
  stage_tile<T, D, LDS>(Vs, vg, p.ld_kv, kLP, p.L, p.d, tid, kWarps * 32);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  if (warp * 16 >= rows_valid) return;  // whole warp beyond the tail (no later block-wide sync)

  float s[kNT][4];
#pragma unroll
  for (int i = 0; i < kNT; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
  warp_qk<T, D, LDS>(s, Qs, warp * 16, Ks, lane);
  warp_softmax(s, p.L, p.scale_log2, lane);
  uint32_t pa[kPK][4];
  pack_afrag<T>(pa, s);
  float o[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  warp_pv<T, D, LDS>(o, pa, Vs, lane);

  T* og = static_cast<T*>(p.o) + ((long long)b * p.N + q0) * p.ld_o + (long long)h * p.d;
  warp_store_rows<T, D, LDS>(o, Qs, warp * 16, og, p.ld_o, rows_valid, p.d, lane);
}

// =============================================================================================================
// backward: grid (ceil(N / rows_per_cta), H, B), 4 warps; the CTA walks its rows 64 at a time, keeps its share of
// dK/dV (all 80 keys x the warp's d-columns) in registers and adds it to the fp32 dK/dV buffers at the end.
// =============================================================================================================
template <typename T, int D>
__global__ void __launch_bounds__(kWarps * 32) attn2_bwd_kernel(const __grid_constant__ AttnParams p) {
  constexpr int LDS = D + 8;
  constexpr int ROWS = kWarps * 16;
  constexpr int DT = D / 8;                            // n8 tiles along d
  constexpr int NW = (DT + kWarps - 1) / kWarps;       // d-tiles owned by one warp (tile = warp + 4*j)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* Ks = reinterpret_cast<T*>(smem_raw);
  T* Vs = Ks + kLP * LDS;
  T* Qs = Vs + kLP * LDS;
  T* dOs = Qs + ROWS * LDS;
  T* dQs = dOs + ROWS * LDS;
  T* Ps = dQs + ROWS * LDS;
  T* dSs = Ps + ROWS * kLDP;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int h = blockIdx.y, b = blockIdx.z;
  const int row_begin = blockIdx.x * p.rows_per_cta;
  const int row_end = min(p.N, row_begin + p.rows_per_cta);

  const T* kg = static_cast<const T*>(p.k) + (long long)b * p.bs_kv + (long long)h * p.d;
  const T* vg = static_cast<const T*>(p.v) + (long long)b * p.bs_kv + (long long)h * p.d;
  stage_tile<T, D, LDS>(Ks, kg, p.ld_kv, kLP, p.L, p.d, tid, kWarps * 32);
  stage_tile<T, D, LDS>(Vs, vg, p.ld_kv, kLP, p.L, p.d, tid, kWarps * 32);

  float dk[kPK][NW][4], dv[kPK][NW][4];
#pragma unroll
  for (int i = 0; i < kPK; ++i)
#pragma unroll
    for (int j = 0; j < NW; ++j) { dk[i][j][0] = dk[i][j][1] = dk[i][j][2] = dk[i][j][3] = 0.f; dv[i][j][0] = dv[i][j][1] = dv[i][j][2] = dv[i][j][3] = 0.f; }

  for (int q0 = row_begin; q0 < row_end; q0 += ROWS) {
    const int rows_valid = min(ROWS, row_end - q0);
    const T* qg = static_cast<const T*>(p.q) + ((long long)b * p.N + q0) * p.ld_q + (long long)h * p.d;
    const T* dog = static_cast<const T*>(p.dout) + ((long long)b * p.N + q0) * p.ld_do + (long long)h * p.d;
    stage_tile<T, D, LDS>(Qs, qg, p.ld_q, ROWS, rows_valid, p.d, tid, kWarps * 32);
    stage_tile<T, D, LDS>(dOs, dog, p.ld_do, ROWS, rows_valid, p.d, tid, kWarps * 32);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    // ---- per-warp: S, P, dP, dS, dQ for rows [warp*16, +16) ----------------------------------------------
    {
      float s[kNT][4], dp[kNT][4];
#pragma unroll
      for (int i = 0; i < kNT; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f; }
      warp_qk<T, D, LDS>(s, Qs, warp * 16, Ks, lane);
      warp_softmax(s, p.L, p.scale_log2, lane);
      warp_qk<T, D, LDS>(dp, dOs, warp * 16, Vs, lane);  // dP = dO V^T
      float d0 = 0.f, d1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < kNT; ++nt) {
        d0 += s[nt][0] * dp[nt][0] + s[nt][1] * dp[nt][1];
        d1 += s[nt][2] * dp[nt][2] + s[nt][3] * dp[nt][3];
      }
      d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
      d1 += __shfl_xor_sync(0xffffffffu, d1, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
      // rows beyond the tail hold zero Q/dO: P is uniform there but dP = 0 -> dS = 0, and P^T dO adds 0 to dV.
      const int r0 = warp * 16 + (lane >> 2);
#pragma unroll
      for (int nt = 0; nt < kNT; ++nt) {
        const int c = nt * 8 + 2 * (lane & 3);
        *reinterpret_cast<uint32_t*>(Ps + r0 * kLDP + c) = pack2<T>(s[nt][0], s[nt][1]);
        *reinterpret_cast<uint32_t*>(Ps + (r0 + 8) * kLDP + c) = pack2<T>(s[nt][2], s[nt][3]);
        dp[nt][0] = s[nt][0] * (dp[nt][0] - d0) * p.scale; dp[nt][1] = s[nt][1] * (dp[nt][1] - d0) * p.scale;
        dp[nt][2] = s[nt][2] * (dp[nt][2] - d1) * p.scale; dp[nt][3] = s[nt][3] * (dp[nt][3] - d1) * p.scale;
        *reinterpret_cast<uint32_t*>(dSs + r0 * kLDP + c) = pack2<T>(dp[nt][0], dp[nt][1]);
        *reinterpret_cast<uint32_t*>(dSs + (r0 + 8) * kLDP + c) = pack2<T>(dp[nt][2], dp[nt][3]);
      }
      uint32_t dsa[kPK][4];
      pack_afrag<T>(dsa, dp);
      float dq[DT][4];
#pragma unroll
      for (int i = 0; i < DT; ++i) { dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f; }
      warp_pv<T, D, LDS>(dq, dsa, Ks, lane);  // dQ = dS K
      T* dqg = static_cast<T*>(p.dq) + ((long long)b * p.N + q0) * p.ld_dq + (long long)h * p.d;
      warp_store_rows<T, D, LDS>(dq, dQs, warp * 16, dqg, p.ld_dq, rows_valid, p.d, lane);
    }
    __syncthreads();  // P / dS of all 64 rows visible

    // ---- dV += P^T dO ; dK += dS^T Q over the 64 rows; warp owns d-tiles {warp, warp+4, ...} ---------------
#pragma unroll
    for (int kq = 0; kq < ROWS / 16; ++kq) {
      uint32_t bdo[NW][2], bq[NW][2];
#pragma unroll
      for (int j = 0; j < NW; ++j) {
        const int dt = warp + kWarps * j;
        if (dt < DT) {
          const int off = (kq * 16 + (lane & 15)) * LDS + dt * 8;
          ldsm_x2_t(smem_u32(dOs + off), bdo[j][0], bdo[j][1]);
          ldsm_x2_t(smem_u32(Qs + off), bq[j][0], bq[j][1]);
        }
      }
#pragma unroll
      for (int mi = 0; mi < kPK; ++mi) {
        uint32_t ap[4], ads[4];
        const int off = (kq * 16 + (lane & 7) + ((lane >> 4) << 3)) * kLDP + mi * 16 + ((lane >> 3) & 1) * 8;
        ldsm_x4_t(smem_u32(Ps + off), ap[0], ap[1], ap[2], ap[3]);
        ldsm_x4_t(smem_u32(dSs + off), ads[0], ads[1], ads[2], ads[3]);
#pragma unroll
        for (int j = 0; j < NW; ++j) {
          if (warp + kWarps * j < DT) {
            mma16816<T>(dv[mi][j], ap, bdo[j][0], bdo[j][1]);
            mma16816<T>(dk[mi][j], ads, bq[j][0], bq[j][1]);
          }
        }
      }
    }
    __syncthreads();  // before the next block's loads overwrite Qs / dOs / Ps / dSs
  }

  // ---- add this CTA's dK / dV share to the fp32 accumulators ------------------------------------------------
  float* dkg = p.dk + (long long)b * p.bs_dkv + (long long)h * p.d;
  float* dvg = p.dv + (long long)b * p.bs_dkv + (long long)h * p.d;
#pragma unroll
  for (int mi = 0; mi < kPK; ++mi)
#pragma unroll
    for (int j = 0; j < NW; ++j) {
      const int dt = warp + kWarps * j;
      const int c = dt * 8 + 2 * (lane & 3);
      if (dt < DT && c < p.d) {
        const int k0 = mi * 16 + (lane >> 2);
        if (k0 < p.L) {
          atomicAdd(reinterpret_cast<float2*>(dkg + (long long)k0 * p.ld_dkv + c), make_float2(dk[mi][j][0], dk[mi][j][1]));
          atomicAdd(reinterpret_cast<float2*>(dvg + (long long)k0 * p.ld_dkv + c), make_float2(dv[mi][j][0], dv[mi][j][1]));
        }
        if (k0 + 8 < p.L) {
          atomicAdd(reinterpret_cast<float2*>(dkg + (long long)(k0 + 8) * p.ld_dkv + c), make_float2(dk[mi][j][2], dk[mi][j][3]));
          atomicAdd(reinterpret_cast<float2*>(dvg + (long long)(k0 + 8) * p.ld_dkv + c), make_float2(dv[mi][j][2], dv[mi][j][3]));
        }
      }
    }
}

template <int D> constexpr size_t fwd_smem() { return (size_t)(2 * kLP + kWarps * 16) * (D + 8) * 2; }
template <int D> constexpr size_t bwd_smem() { return (size_t)(2 * kLP + 3 * kWarps * 16) * (D + 8) * 2 + (size_t)2 * kWarps * 16 * kLDP * 2; }

template <typename T, int D>
int launch_fwd(const AttnParams& p, cudaStream_t st) {
  static std::atomic<unsigned long long> attr_done{0};
  if (int rc = ensure_dynamic_smem(attn2_fwd_kernel<T, D>, (int)fwd_smem<D>(), attr_done)) return rc;
  dim3 grid((p.N + kWarps * 16 - 1) / (kWarps * 16), p.H, p.B);
  attn2_fwd_kernel<T, D><<<grid, kWarps * 32, fwd_smem<D>(), st>>>(p);
  PEA_LAUNCH_CHECK();
  return PEA_OK;
}
template <typename T, int D>
int launch_bwd(const AttnParams& p, cudaStream_t st) {
  static std::atomic<unsigned long long> attr_done{0};
  if (int rc = ensure_dynamic_smem(attn2_bwd_kernel<T, D>, (int)bwd_smem<D>(), attr_done)) return rc;
  dim3 grid((p.N + p.rows_per_cta - 1) / p.rows_per_cta, p.H, p.B);
  attn2_bwd_kernel<T, D><<<grid, kWarps * 32, bwd_smem<D>(), st>>>(p);
  PEA_LAUNCH_CHECK();
  return PEA_OK;
}

int padded_head_dim(int d) { return d <= 48 ? 48 : d <= 64 ? 64 : d <= 80 ? 80 : d <= 160 ? 160 : 0; }

int validate(const pea_attn2_desc* a, bool bwd, AttnParams& p) {
  PEA_CHECK_ARG(a, "attn2: null descriptor");
  PEA_CHECK_ARG(a->B >= 1 && a->N >= 1 && a->H >= 1, "attn2: B, N, H must be >= 1");
  PEA_CHECK_ARG(a->L >= 1 && a->L <= kLP, "attn2: L=%d outside [1,%d] (the key tile is a single %d-wide tile)", a->L, kLP, kLP);
  PEA_CHECK_ARG(a->d >= 8 && a->d % 8 == 0 && padded_head_dim(a->d) != 0, "attn2: head_dim %d unsupported (need multiple of 8, <= 160)", a->d);
  PEA_CHECK_ARG(a->dtype == PEA_BF16 || a->dtype == PEA_F16, "attn2: dtype must be bf16 or fp16");
  PEA_CHECK_ARG(a->d_q && a->d_k && a->d_v, "attn2: null q/k/v");
  auto al16 = [](const void* x) { return (reinterpret_cast<uintptr_t>(x) & 15) == 0; };
  PEA_CHECK_ARG(al16(a->d_q) && al16(a->d_k) && al16(a->d_v) && a->ld_q % 8 == 0 && a->ld_kv % 8 == 0 && a->bs_kv % 8 == 0,
                "attn2: q/k/v must be 16-byte aligned with row strides that are multiples of 8 elements");
  p.B = a->B; p.N = a->N; p.L = a->L; p.H = a->H; p.d = a->d;
  p.scale = a->scale; p.scale_log2 = a->scale * 1.4426950408889634f;
  p.q = a->d_q; p.ld_q = a->ld_q; p.k = a->d_k; p.v = a->d_v; p.ld_kv = a->ld_kv; p.bs_kv = a->bs_kv;
  if (!bwd) {
    PEA_CHECK_ARG(a->d_o && al16(a->d_o) && a->ld_o % 8 == 0, "attn2_fwd: o must be non-null, 16-byte aligned, ld_o %% 8 == 0");
    p.o = a->d_o; p.ld_o = a->ld_o;
  } else {
    PEA_CHECK_ARG(a->d_do && a->d_dq && a->d_dk && a->d_dv, "attn2_bwd: null do/dq/dk/dv");
    PEA_CHECK_ARG(al16(a->d_do) && al16(a->d_dq) && a->ld_do % 8 == 0 && a->ld_dq % 8 == 0, "attn2_bwd: do/dq alignment");
    PEA_CHECK_ARG((reinterpret_cast<uintptr_t>(a->d_dk) & 7) == 0 && (reinterpret_cast<uintptr_t>(a->d_dv) & 7) == 0 &&
                      a->ld_dkv % 2 == 0 && a->bs_dkv % 2 == 0, "attn2_bwd: dk/dv must be 8-byte aligned with even strides");
    p.dout = a->d_do; p.ld_do = a->ld_do; p.dq = a->d_dq; p.ld_dq = a->ld_dq;
    p.dk = a->d_dk; p.dv = a->d_dv; p.ld_dkv = a->ld_dkv; p.bs_dkv = a->bs_dkv;
    // rows per CTA: as few CTAs per (b, h) as still fill the machine ~2x (fewer dK/dV atomics), at least 64 rows.
    const long long target_ctas = (long long)num_sms() * 4;
    long long per_bh = (target_ctas + (long long)a->B * a->H - 1) / ((long long)a->B * a->H);
    if (per_bh < 1) per_bh = 1;
    long long rows = (a->N + per_bh - 1) / per_bh;
    rows = ((rows + 63) / 64) * 64;
    p.rows_per_cta = (int)rows;
  }
  return PEA_OK;
}

}  // namespace
}  // namespace pea

namespace pea {
int attn2_fwd_tc(const pea_attn2_desc* a, cudaStream_t st);   // attn2_tc.cu
bool attn2_fwd_tc_eligible(const pea_attn2_desc* a);
int attn2_bwd_tc(const pea_attn2_desc* a, cudaStream_t st);   // attn2_bwd_tc.cu
bool attn2_bwd_tc_eligible(const pea_attn2_desc* a);
}  // namespace pea

using namespace pea;

// PEA_ATTN2_IMPL=mma forces the warp-MMA kernels (A/B measurements); default: tcgen05 forward when head_dim == 64.
static bool use_tc_path() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PEA_ATTN2_IMPL");
    v = (e && e[0] == 'm') ? 0 : 1;
  }
  return v == 1;
}

#define PEA_DISPATCH_ATTN(LAUNCH)                                                                           \
  do {                                                                                                       \
    const int D = padded_head_dim(p.d);                                                                      \
    if (desc->dtype == PEA_BF16) {                                                                           \
      switch (D) {                                                                                           \
        case 48: return LAUNCH<__nv_bfloat16, 48>(p, st);                                                    \
        case 64: return LAUNCH<__nv_bfloat16, 64>(p, st);                                                    \
        case 80: return LAUNCH<__nv_bfloat16, 80>(p, st);                                                    \
        case 160: return LAUNCH<__nv_bfloat16, 160>(p, st);                                                  \
      }                                                                                                      \
    } else {                                                                                                 \
      switch (D) {                                                                                           \
        case 48: return LAUNCH<__half, 48>(p, st);                                                           \
        case 64: return LAUNCH<__half, 64>(p, st);                                                           \
        case 80: return LAUNCH<__half, 80>(p, st);                                                           \
        case 160: return LAUNCH<__half, 160>(p, st);                                                         \
      }                                                                                                      \
    }                                                                                                        \
    pea::set_error("attn2: no kernel for head_dim %d", p.d);                                                 \
    return PEA_EUNSUPPORTED;                                                                                 \
  } while (0)

extern "C" int pea_attn2_fwd(const pea_attn2_desc* desc, void* stream) {
  AttnParams p{};
  int rc = validate(desc, false, p);
  if (rc != PEA_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (use_tc_path() && attn2_fwd_tc_eligible(desc)) return attn2_fwd_tc(desc, st);
  PEA_DISPATCH_ATTN(launch_fwd);
}

extern "C" int pea_attn2_bwd(const pea_attn2_desc* desc, void* stream) {
  AttnParams p{};
  int rc = validate(desc, true, p);
  if (rc != PEA_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (use_tc_path() && attn2_bwd_tc_eligible(desc)) return attn2_bwd_tc(desc, st);
  PEA_DISPATCH_ATTN(launch_bwd);
}
