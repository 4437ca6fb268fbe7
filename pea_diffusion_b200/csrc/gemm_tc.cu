// gemm_tc.cu -- tcgen05 / TMEM / TMA GEMM with fused epilogues (SURVEY.md kernels K1, K2 and the dContext GEMM).
//
//   C[m, n] = sum_k A(m, k) * B(n, k)
//
// Used for  * the PEA adapter's Linear chain forward and backward (train_sdxl_zh.py:58-67 and its autograd graph):
//             bias, exact-erf GELU, GELU', the mean-pool gradient broadcast and the bf16 operand split of the
//             fp32-accurate mode all live in the epilogue;
//           * attn2's to_k / to_v projections of the context (diffusers AttnProcessor2_0) and dContext = dK Wk + dV Wv.
//
// Structure (persistent CTAs, 128 x bn output tiles, bn a multiple of 16 chosen per problem against the wave count, 320 threads):
//   warp 0      TMA producer: 128B-swizzled [128 x 64] / [BN x 64] operand tiles, `stages`-deep mbarrier ring that runs
//               across tile boundaries
//   warp 1      TMEM allocator + MMA issuer: one elected lane issues tcgen05.mma (M128 N{128,256} K16, kind::f16, fp32
//               accumulate); accumulators are DOUBLE BUFFERED in TMEM (2 x BN columns) so tile i+1's main loop runs
//               while the epilogue drains tile i; tcgen05.commit releases smem stages and signals the epilogue
//   warps 2..9  epilogue: tcgen05.ld (each warp its 32-lane TMEM quadrant, thread = output row; two warps per quadrant, half
//               the tile's columns each), fused math, stores
// Operand majors: K-major tiles come from one {64, 128} TMA box; MN-major tiles (A^T / B^T stored row-major) from two
// {64, 64} boxes, consumed through MN-major UMMA descriptors -- so dX = dY W and dW = dY^T X need no transposes.
// fp32-accurate mode: operands are sums of up to 3 bf16 planes; per K step the kernel issues the cross products
// A[i] B[j], i + j < S.  All planes of a stage share one mbarrier transaction.
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <unordered_map>

#include "tc05.cuh"

namespace pea {
namespace {

constexpr int BM = 128, BK = 64;
constexpr int kTileBytes = BM * BK * 2;  // 16 KiB per 128-row operand plane per stage
constexpr int kMaxStages = 8;
constexpr int kThreads = 192;        // gemm_tc2_kernel (CTA pairs): producer, MMA issuer, 4 epilogue warps
constexpr int kThreads1 = 320;       // gemm_tc_kernel: producer, MMA issuer, 8 epilogue warps (two per TMEM lane quadrant, half the
                                     // tile's columns each: a one-wave GEMM's epilogue is fully exposed, so its latency is halved)
constexpr size_t kSmemBudget = 208 * 1024;

struct GemmParams {
  CUtensorMap tmA[3];
  CUtensorMap tmB[3];
  int M, N, K;
  int sa, sb;
  int a_mn, b_mn;
  int kb_total, kb_per_split;
  int m_tiles, n_tiles, splits, total_tiles;
  int stages;
  int bn;                    // output tile width (multiple of 16, 32..256), chosen per problem against the wave count
  uint32_t b_tile_bytes;     // one B plane of one stage: bn x 128 B (K-major) or ceil(bn / 64) x 8 KiB (MN-major slabs)
  uint32_t idesc;
  float alpha;
  const float* bias;
  const void* aux; int aux_dtype; long long ldaux;
  const float* rowadd; int rowadd_group; float rowadd_scale;
  int act;
  void* out0; int out0_dtype; long long ldo0; int out0_mode;  // 0 store, 1 +=, 2 atomic add
  void* out1[3]; int out1_nsplit; int out1_dtype; long long ldo1;
  int pdl;
  int epi_plain;   // out0 = alpha * acc (+ bias), plain vectorised store: the specialised epilogue below
};

// ---- epilogue helpers: 8 consecutive columns of one row -----------------------------------------------------------
__device__ __forceinline__ void load8(const void* base, int dtype, long long off, bool vec, int valid, float* f) {
  if (dtype == PEA_F32) {
    const float* p = static_cast<const float*>(base) + off;
    if (vec) {
      const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
      f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = j < valid ? p[j] : 0.f;
    }
  } else if (dtype == PEA_F16) {
    const __half* p = static_cast<const __half*>(base) + off;
    if (vec) {
      const uint4 u = *reinterpret_cast<const uint4*>(p);
      float2 a = unpack2<__half>(u.x), b = unpack2<__half>(u.y), c = unpack2<__half>(u.z), d = unpack2<__half>(u.w);
      f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = j < valid ? __half2float(p[j]) : 0.f;
    }
  } else {
    const __nv_bfloat16* p = static_cast<const __nv_bfloat16*>(base) + off;
    if (vec) {
      const uint4 u = *reinterpret_cast<const uint4*>(p);
      float2 a = unpack2<__nv_bfloat16>(u.x), b = unpack2<__nv_bfloat16>(u.y), c = unpack2<__nv_bfloat16>(u.z), d = unpack2<__nv_bfloat16>(u.w);
      f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = j < valid ? __bfloat162float(p[j]) : 0.f;
    }
  }
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}

__device__ __forceinline__ void store8(void* base, int dtype, long long off, bool vec, int valid, const float* f, int mode) {
  if (dtype == PEA_F32) {
    float* p = static_cast<float*>(base) + off;
    if (mode == 2) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < valid) atomicAdd(p + j, f[j]);
    } else if (vec) {
      float4 a = make_float4(f[0], f[1], f[2], f[3]), b = make_float4(f[4], f[5], f[6], f[7]);
      if (mode == 1) {
        const float4 oa = *reinterpret_cast<const float4*>(p), ob = *reinterpret_cast<const float4*>(p + 4);
        a.x += oa.x; a.y += oa.y; a.z += oa.z; a.w += oa.w; b.x += ob.x; b.y += ob.y; b.z += ob.z; b.w += ob.w;
      }
      *reinterpret_cast<float4*>(p) = a;
      *reinterpret_cast<float4*>(p + 4) = b;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < valid) p[j] = (mode == 1 ? p[j] : 0.f) + f[j];
    }
  } else if (dtype == PEA_F16) {
    __half* p = static_cast<__half*>(base) + off;
    float g[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = f[j];
    if (mode == 1) {
      float o[8];
      load8(base, dtype, off, vec, valid, o);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] += o[j];
    }
    if (vec) {
      *reinterpret_cast<uint4*>(p) = make_uint4(pack2<__half>(g[0], g[1]), pack2<__half>(g[2], g[3]), pack2<__half>(g[4], g[5]), pack2<__half>(g[6], g[7]));
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < valid) p[j] = __float2half_rn(g[j]);
    }
  } else {
    __nv_bfloat16* p = static_cast<__nv_bfloat16*>(base) + off;
    float g[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = f[j];
    if (mode == 1) {
      float o[8];
      load8(base, dtype, off, vec, valid, o);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] += o[j];
    }
    if (vec) {
      *reinterpret_cast<uint4*>(p) = make_uint4(pack2<__nv_bfloat16>(g[0], g[1]), pack2<__nv_bfloat16>(g[2], g[3]),
                                                pack2<__nv_bfloat16>(g[4], g[5]), pack2<__nv_bfloat16>(g[6], g[7]));
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < valid) p[j] = __float2bfloat16_rn(g[j]);
    }
  }
}

__device__ __forceinline__ float round16(float v, int dtype) {
  return dtype == PEA_F16 ? __half2float(__float2half_rn(v)) : __bfloat162float(__float2bfloat16_rn(v));
}

// Specialised epilogue for the common case out0 = alpha * acc (+ bias[n]) stored once (any dtype, N % 8 == 0, aligned): 32
// accumulator columns per TMEM load, the next load in flight while the current columns are converted and stored with 128-bit
// stores.  The general body below walks ~100 mostly-branch instructions per 8 columns in a single warp per scheduler:
// 7 600 cycles for a 128 x 128 tile (profiles/r2_gemm_timeline.txt) -- as long as the whole K = 1024 main loop of the
// one-wave adapter GEMMs.  This one takes ~1 500.
// 32 accumulator columns of one row: alpha, bias, convert, four 128-bit (two for 16-bit outputs: 8 columns = 16 B) stores.
template <int DT>
__device__ __forceinline__ void plain_store32(const GemmParams& p, const uint32_t (&v)[32], int row, int nb, int n_end, float alpha) {
#pragma unroll
  for (int g8 = 0; g8 < 4; ++g8) {
    const int n = nb + g8 * 8;
    if (n < n_end) {        // N and bn are multiples of 8: a group of 8 columns is either all in or all out
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = alpha * __uint_as_float(v[g8 * 8 + j]);
      if (p.bias) {
        const float4 b0 = *reinterpret_cast<const float4*>(p.bias + n), b1 = *reinterpret_cast<const float4*>(p.bias + n + 4);
        f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w; f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
      }
      const long long off = (long long)row * p.ldo0 + n;
      if constexpr (DT == PEA_F32) {
        float* o = static_cast<float*>(p.out0) + off;
        *reinterpret_cast<float4*>(o) = make_float4(f[0], f[1], f[2], f[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(f[4], f[5], f[6], f[7]);
      } else if constexpr (DT == PEA_BF16) {
        *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out0) + off) =
            make_uint4(pack2<__nv_bfloat16>(f[0], f[1]), pack2<__nv_bfloat16>(f[2], f[3]), pack2<__nv_bfloat16>(f[4], f[5]), pack2<__nv_bfloat16>(f[6], f[7]));
      } else {
        *reinterpret_cast<uint4*>(static_cast<__half*>(p.out0) + off) =
            make_uint4(pack2<__half>(f[0], f[1]), pack2<__half>(f[2], f[3]), pack2<__half>(f[4], f[5]), pack2<__half>(f[6], f[7]));
      }
    }
  }
}

template <int DT>
__device__ __forceinline__ void epilogue_tile_plain_dt(const GemmParams& p, uint32_t tacc, int row, int n0, int bn) {
  const bool row_ok = row < p.M;
  const float alpha = p.alpha;
  const int n_end = min(p.N, n0 + bn);     // columns past the tile belong to the next tile (or to nobody)
  uint32_t va[32], vb[32];
  tc::tmem_ld32(tacc, va);
  // two 32-column chunks per (rolled) iteration, ping-pong buffers: the next chunk's TMEM load is always in flight.  Rolled so
  // that the executed code stays a few KB (a fully unrolled 256-column body is 40 KB: it thrashes the instruction cache).
#pragma unroll 1
  for (int c = 0; c * 32 < bn; c += 2) {
    tc::tmem_ld_wait();
    if ((c + 1) * 32 < bn) tc::tmem_ld32(tacc + (uint32_t)((c + 1) * 32), vb);     // warp-uniform
    if (row_ok) plain_store32<DT>(p, va, row, n0 + c * 32, n_end, alpha);
    if ((c + 1) * 32 < bn) {
      tc::tmem_ld_wait();
      if ((c + 2) * 32 < bn) tc::tmem_ld32(tacc + (uint32_t)((c + 2) * 32), va);
      if (row_ok) plain_store32<DT>(p, vb, row, n0 + (c + 1) * 32, n_end, alpha);
    }
  }
}

__device__ __forceinline__ void epilogue_tile_plain(const GemmParams& p, uint32_t tacc, int row, int n0, int bn) {
  if (p.out0_dtype == PEA_F32) epilogue_tile_plain_dt<PEA_F32>(p, tacc, row, n0, bn);
  else if (p.out0_dtype == PEA_BF16) epilogue_tile_plain_dt<PEA_BF16>(p, tacc, row, n0, bn);
  else epilogue_tile_plain_dt<PEA_F16>(p, tacc, row, n0, bn);
}

// Epilogue of one 128-row x BN-column accumulator tile for one thread (= one output row): TMEM -> registers -> fused math
// -> global.  `tacc` = TMEM address of the accumulator's column 0 in this warp's lane quadrant.
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, uint32_t tacc, int row, int n0, int bn, bool al0, bool al1, bool alx) {
  if (p.epi_plain) { epilogue_tile_plain(p, tacc, row, n0, bn); return; }
  const bool row_ok = row < p.M;
  const float* radd = (p.rowadd && row_ok) ? p.rowadd + (long long)(row / p.rowadd_group) * p.N : nullptr;
  // ONE rolled copy of the 8-column body.  With every dtype / mode variant inlined it is ~1.5 k instructions; unrolled four
  // times it made the kernel 112 KB of code, more than the 32 KB instruction cache, and a one-tile-per-CTA GEMM (the adapter
  // shapes) spent most of its 22 us fetching epilogue instructions (ncu r1: 42 k cycles for 16 k-blocks).  The TMEM load of
  // the next 8 columns is issued before the current ones are processed, so its latency stays hidden.
  uint32_t v[8];
  tmem_ld8(tacc, v);
  const int ngroups = bn >> 3;
#pragma unroll 1
  for (int c = 0; c < ngroups; ++c) {
    tc::tmem_ld_wait();
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = p.alpha * __uint_as_float(v[j]);
    if (c + 1 < ngroups) tmem_ld8(tacc + (uint32_t)((c + 1) * 8), v);   // warp-uniform: all 32 lanes issue it
    const int n = n0 + c * 8;
    if (row_ok && n < p.N) {
      const int valid = min(8, p.N - n);
      const bool full = valid == 8;
      if (p.bias) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j < valid) f[j] += p.bias[n + j];
      }
      if (p.aux) {
        float a[8];
        load8(p.aux, p.aux_dtype, (long long)row * p.ldaux + n, full && alx, valid, a);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] *= gelu_exact_grad(a[j]);
      }
      if (radd) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j < valid) f[j] = fmaf(radd[n + j], p.rowadd_scale, f[j]);
      }
      if (p.out0) store8(p.out0, p.out0_dtype, (long long)row * p.ldo0 + n, full && al0, valid, f, p.out0_mode);
      if (p.out1_nsplit > 0) {
        if (p.act) {
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = gelu_exact(f[j]);
        }
#pragma unroll 1
        for (int s = 0; s < p.out1_nsplit; ++s) {
          float r[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) { r[j] = round16(f[j], p.out1_dtype); f[j] -= r[j]; }
          store8(p.out1[s], p.out1_dtype, (long long)row * p.ldo1 + n, full && al1, valid, r, 0);
        }
      }
    }
  }
}

__device__ __forceinline__ void epilogue_alignment(const GemmParams& p, bool& al0, bool& al1, bool& alx) {
  al0 = p.out0 && (p.ldo0 % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.out0) & (p.out0_dtype == PEA_F32 ? 31 : 15)) == 0);
  al1 = p.out1_nsplit > 0 && (p.ldo1 % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.out1[0]) & 15) == 0) &&
        (p.out1_nsplit < 2 || (reinterpret_cast<uintptr_t>(p.out1[1]) & 15) == 0) &&
        (p.out1_nsplit < 3 || (reinterpret_cast<uintptr_t>(p.out1[2]) & 15) == 0);
  alx = p.aux && (p.ldaux % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.aux) & (p.aux_dtype == PEA_F32 ? 31 : 15)) == 0);
}

// Optional event timeline of CTA 0 (build with -DPEA_TRACE, read with tools/trace_gemm.py).  No-op in the product build.
#ifdef PEA_TRACE
__device__ unsigned long long g_trace_gemm[64];
#define GTRACE(ev) do { if (blockIdx.x == 0 && lane == 0) g_trace_gemm[ev] = clock64(); } while (0)
#else
#define GTRACE(ev) do { } while (0)
#endif

// SA / SB: operand plane counts; A_MN / B_MN: operand majors.  Everything the issue loops touch is a compile-time
// constant or a warp-uniform value: the whole producer / MMA warp runs the loop (uniform control flow keeps address
// and descriptor arithmetic on the uniform datapath) and only the asynchronous instruction itself is issued by one
// elected lane.  A single-lane loop with runtime plane / major branches measured ~430 cycles between MMAs (the
// 128x128x16 MMA itself takes 64): r1 ncu capture profiles/gemm_r1_before.txt.
template <int SA, int SB, int A_MN, int B_MN>
__global__ void __launch_bounds__(kThreads1, 1) gemm_tc_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_slot;

  constexpr int S = SA > SB ? SA : SB;
  const int bn = p.bn;
  const uint32_t kBTile = p.b_tile_bytes;                                  // B plane bytes per stage
  const uint32_t kStageBytes = (uint32_t)SA * kTileBytes + (uint32_t)SB * kBTile;
  constexpr uint32_t kTmemCols = 512;                                      // two accumulators of up to 256 columns
  constexpr uint32_t kAccStride = 256;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;  // 128B swizzle atoms need 1024-byte alignment
  const int stages = p.stages;
  if (warp == 0) GTRACE(0);   // kernel entry
  // tile index -> (split, n tile, m tile), m fastest: CTAs that run together share B (weight) tiles through L2
  auto tile_coords = [&](int tile, int& m0, int& n0, int& kb_begin, int& nkb) {
    const int mi = tile % p.m_tiles;
    const int rest = tile / p.m_tiles;
    const int ni = rest % p.n_tiles, z = rest / p.n_tiles;
    m0 = mi * BM; n0 = ni * bn;
    kb_begin = z * p.kb_per_split;
    nkb = min(p.kb_total, kb_begin + p.kb_per_split) - kb_begin;
  };

  if (warp == 0 && lane == 0) {
#pragma unroll
    for (int i = 0; i < SA; ++i) tc::prefetch_tmap(&p.tmA[i]);
#pragma unroll
    for (int i = 0; i < SB; ++i) tc::prefetch_tmap(&p.tmB[i]);
    for (int s = 0; s < stages; ++s) {
      tc::mbar_init(smem_u32(&full_bar[s]), 1);
      tc::mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    tc::mbar_init(smem_u32(&acc_full[0]), 1); tc::mbar_init(smem_u32(&acc_full[1]), 1);
    tc::mbar_init(smem_u32(&acc_empty[0]), 8); tc::mbar_init(smem_u32(&acc_empty[1]), 8);
    tc::fence_mbar_init();
  }
  if (warp == 1) tc::tmem_alloc(smem_u32(&tmem_base_slot), kTmemCols);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_wait();                 // PEA_LAUNCH_PDL: the prologue above overlapped the preceding kernel's tail
  pdl_launch_dependents();
  if (warp == 0) GTRACE(1);   // prologue done

  if (warp == 0) {
    // ===== TMA producer (whole warp loops; one elected lane issues) =====
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int m0, n0, kb_begin, nkb;
      tile_coords(tile, m0, n0, kb_begin, nkb);
      for (int it = 0; it < nkb; ++it) {
        tc::mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
        if (tc::elect_one()) {
          const uint32_t fb = smem_u32(&full_bar[s]);
          tc::mbar_expect_tx(fb, kStageBytes);
          const int k0 = (kb_begin + it) * BK;
          const uint32_t dst = tiles + (uint32_t)s * kStageBytes;
#pragma unroll
          for (int i = 0; i < SA; ++i) {
            const uint32_t d = dst + (uint32_t)i * kTileBytes;
            if (!A_MN) tc::tma_load_2d(d, &p.tmA[i], fb, k0, m0);
            else { tc::tma_load_2d(d, &p.tmA[i], fb, m0, k0); tc::tma_load_2d(d + 8192, &p.tmA[i], fb, m0 + 64, k0); }
          }
#pragma unroll
          for (int j = 0; j < SB; ++j) {
            const uint32_t d = dst + (uint32_t)SA * kTileBytes + (uint32_t)j * kBTile;
            if (!B_MN) tc::tma_load_2d(d, &p.tmB[j], fb, k0, n0);
            else {
#pragma unroll
              for (int c = 0; c < 4; ++c)
                if (c * 64 < bn) tc::tma_load_2d(d + c * 8192, &p.tmB[j], fb, n0 + c * 64, k0);
            }
          }
        }
        __syncwarp();
        if (++s == stages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (whole warp loops; one elected lane issues tcgen05.mma / tcgen05.commit) =====
    // per K=16 step: K-major advances 32 B inside the swizzle atom; MN-major advances 16 k-rows = 2048 B
    constexpr uint32_t kAStep = A_MN ? 2048u : 32u, kBStep = B_MN ? 2048u : 32u;
    constexpr uint32_t kALbo = A_MN ? (uint32_t)BK * 128u : 16u, kBLbo = B_MN ? (uint32_t)BK * 128u : 16u;
    // descriptor = constant high word | (lbo field) | 14-bit (address >> 4); only the address part changes
    constexpr uint64_t kADescHi = ((uint64_t)(kALbo >> 4) << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
    constexpr uint64_t kBDescHi = ((uint64_t)(kBLbo >> 4) << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
    const uint32_t idesc = p.idesc;
    int s = 0;
    uint32_t ph = 0;
    int local = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++local) {
      int m0, n0, kb_begin, nkb;
      tile_coords(tile, m0, n0, kb_begin, nkb);
      const int as = local & 1;
      const uint32_t tmem_acc = tmem_base + (uint32_t)as * kAccStride;
      tc::mbar_wait(smem_u32(&acc_empty[as]), ((uint32_t)(local >> 1) & 1u) ^ 1u);   // epilogue drained this accumulator
      tc::fence_after_sync();
      for (int it = 0; it < nkb; ++it) {
        tc::mbar_wait(smem_u32(&full_bar[s]), ph);
        if (local == 0 && it == 0) GTRACE(2);              // first stage landed
        if (local == 0 && it == nkb - 1) GTRACE(3);        // last stage of the first tile landed
        tc::fence_after_sync();
        if (tc::elect_one()) {
          const uint32_t a_base = (tiles + (uint32_t)s * kStageBytes) >> 4;  // smem < 256 KiB: fits the 14-bit field
          const uint32_t b_base = a_base + (uint32_t)SA * (kTileBytes >> 4);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
#pragma unroll
            for (int i = 0; i < SA; ++i) {
#pragma unroll
              for (int j = 0; j < SB; ++j) {
                if (i + j < S) {
                  const uint64_t da = kADescHi | (uint64_t)(a_base + (uint32_t)((i * kTileBytes + k * kAStep) >> 4));
                  const uint64_t db = kBDescHi | (uint64_t)(b_base + (uint32_t)((j * kBTile + k * kBStep) >> 4));
                  tc::mma_f16_ss(tmem_acc, da, db, idesc, (k == 0 && i == 0 && j == 0) ? (uint32_t)(it > 0) : 1u);
                }
              }
            }
          }
          tc::mma_commit(smem_u32(&empty_bar[s]));                      // stage reusable once these MMAs have read it
          if (it == nkb - 1) tc::mma_commit(smem_u32(&acc_full[as]));   // accumulator complete
        }
        __syncwarp();
        if (++s == stages) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    // ===== epilogue: warp w owns TMEM lanes [32*(w%4), +32) = output rows m0 + 32*(w%4) + lane =====
    const int q = warp & 3;
    bool al0, al1, alx;
    epilogue_alignment(p, al0, al1, alx);
    int local = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++local) {
      int m0, n0, kb_begin, nkb;
      tile_coords(tile, m0, n0, kb_begin, nkb);
      const int as = local & 1;
      tc::mbar_wait(smem_u32(&acc_full[as]), (uint32_t)(local >> 1) & 1u);
      if (local == 0 && warp == 2) GTRACE(4);              // first accumulator complete
      tc::fence_after_sync();
      // warps 2-5 take the tile's first bn / 2 columns, warps 6-9 the rest (bn is a multiple of 16, so both halves are
      // multiples of 8 columns)
      const int half = (warp - 2) >> 2, hw = bn >> 1;
      epilogue_tile(p, tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)as * kAccStride + (uint32_t)(half * hw), m0 + q * 32 + lane,
                    n0 + half * hw, hw, al0, al1, alx);
      if (local == 0 && warp == 2) GTRACE(5);              // first tile's epilogue done (this warp)
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(smem_u32(&acc_empty[as]));   // this accumulator may be overwritten (tile local + 2)
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) GTRACE(6);   // all roles done
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, kTmemCols);
  }
}

#ifdef PEA_TRACE
}  // namespace
}  // namespace pea
extern "C" int pea_debug_gemm_trace(unsigned long long* host, int n) {
  return (int)cudaMemcpyFromSymbol(host, pea::g_trace_gemm, sizeof(unsigned long long) * n);
}
namespace pea {
namespace {
#endif

// ---------------------------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): a cluster of two CTAs on adjacent SMs computes one 256 x 256 tile.  Each CTA stages its
// own 128 rows of A and its own 128-row half of the B tile; the leader's single thread issues M256 N256 K16 MMAs that read
// both CTAs' shared memory.  The single-CTA kernel is bound by shared-memory bandwidth -- every operand byte is written by
// TMA and read by the MMA through the same 128 B/clk: 48 KB + 48 KB per 512 cycles of math = 192 B/clk -> 67 % duty, while
// the MMA alone reaches 98 % (tools/mma_probe.cu) -- and the pair moves only 32 KB + 32 KB per SM for the same math.
// Barrier protocol: TMA loads of BOTH CTAs credit the LEADER's full barrier; tcgen05.commit multicasts to the empty /
// accumulator-full barriers of both CTAs; the peer's epilogue warps release the accumulator on the leader's barrier through
// a shared::cluster arrive.  Single-plane operands only (the throughput mode).
template <int A_MN, int B_MN>
__global__ void __launch_bounds__(kThreads, 1) gemm_tc2_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_slot;

  constexpr int BN = 256;                                   // per pair; each CTA stages BN / 2 rows of B
  constexpr uint32_t kBHalf = (uint32_t)(BN / 2) * BK * 2;   // 16 KiB
  constexpr uint32_t kStageBytes = kTileBytes + kBHalf;      // per CTA
  constexpr uint32_t kTmemCols = 2 * BN;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = tc::cluster_ctarank();
  const bool leader = rank == 0;
  const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int stages = p.stages;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  auto tile_coords = [&](int tile, int& m0, int& n0) {   // m_tiles counts 256-row pair tiles; m fastest
    const int mi = tile % p.m_tiles, ni = tile / p.m_tiles;
    m0 = mi * 256; n0 = ni * BN;
  };

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&p.tmA[0]);
    tc::prefetch_tmap(&p.tmB[0]);
    for (int s = 0; s < stages; ++s) {
      tc::mbar_init(smem_u32(&full_bar[s]), 1);
      tc::mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    tc::mbar_init(smem_u32(&acc_full[0]), 1); tc::mbar_init(smem_u32(&acc_full[1]), 1);
    tc::mbar_init(smem_u32(&acc_empty[0]), 8); tc::mbar_init(smem_u32(&acc_empty[1]), 8);   // 4 epilogue warps x 2 CTAs
    tc::fence_mbar_init();
  }
  if (warp == 1) tc::tmem_alloc_cg2(smem_u32(&tmem_base_slot), kTmemCols);
  tc::fence_before_sync();
  __syncthreads();
  tc::cluster_sync();            // the peer's barriers exist before anything arrives on them
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own A rows, own half of B; bytes credited to the leader's full barrier =====
    int s = 0;
    uint32_t ph = 0;
    for (int tile = pair; tile < p.total_tiles; tile += npairs) {
      int m0, n0;
      tile_coords(tile, m0, n0);
      const int ma = m0 + (int)rank * 128, nb = n0 + (int)rank * (BN / 2);
      for (int it = 0; it < p.kb_total; ++it) {
        tc::mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
        if (tc::elect_one()) {
          if (leader) tc::mbar_expect_tx(smem_u32(&full_bar[s]), 2 * kStageBytes);
          const uint32_t fb = tc::mapa(smem_u32(&full_bar[s]), 0);   // the leader's barrier, from either CTA
          const int k0 = it * BK;
          const uint32_t dst = tiles + (uint32_t)s * kStageBytes;
          if (!A_MN) tc::tma_load_2d_cg2(dst, &p.tmA[0], fb, k0, ma);
          else { tc::tma_load_2d_cg2(dst, &p.tmA[0], fb, ma, k0); tc::tma_load_2d_cg2(dst + 8192, &p.tmA[0], fb, ma + 64, k0); }
          const uint32_t db = dst + kTileBytes;
          if (!B_MN) tc::tma_load_2d_cg2(db, &p.tmB[0], fb, k0, nb);
          else { tc::tma_load_2d_cg2(db, &p.tmB[0], fb, nb, k0); tc::tma_load_2d_cg2(db + 8192, &p.tmB[0], fb, nb + 64, k0); }
        }
        __syncwarp();
        if (++s == stages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: leader CTA only =====
    if (leader) {
      constexpr uint32_t kAStep = A_MN ? 2048u : 32u, kBStep = B_MN ? 2048u : 32u;
      constexpr uint32_t kALbo = A_MN ? (uint32_t)BK * 128u : 16u, kBLbo = B_MN ? (uint32_t)BK * 128u : 16u;
      constexpr uint64_t kADescHi = ((uint64_t)(kALbo >> 4) << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
      constexpr uint64_t kBDescHi = ((uint64_t)(kBLbo >> 4) << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
      const uint32_t idesc = p.idesc;
      int s = 0;
      uint32_t ph = 0;
      int local = 0;
      for (int tile = pair; tile < p.total_tiles; tile += npairs, ++local) {
        const int as = local & 1;
        const uint32_t tmem_acc = tmem_base + (uint32_t)as * BN;
        tc::mbar_wait(smem_u32(&acc_empty[as]), ((uint32_t)(local >> 1) & 1u) ^ 1u);   // both CTAs' epilogues drained it
        tc::fence_after_sync();
        for (int it = 0; it < p.kb_total; ++it) {
          tc::mbar_wait(smem_u32(&full_bar[s]), ph);
          tc::fence_after_sync();
          if (tc::elect_one()) {
            const uint32_t a_base = (tiles + (uint32_t)s * kStageBytes) >> 4;
            const uint32_t b_base = a_base + (kTileBytes >> 4);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t da = kADescHi | (uint64_t)(a_base + (uint32_t)((k * kAStep) >> 4));
              const uint64_t db = kBDescHi | (uint64_t)(b_base + (uint32_t)((k * kBStep) >> 4));
              tc::mma_f16_ss_cg2(tmem_acc, da, db, idesc, (k == 0) ? (uint32_t)(it > 0) : 1u);
            }
            tc::mma_commit_cg2(smem_u32(&empty_bar[s]));                          // stage reusable in BOTH CTAs
            if (it == p.kb_total - 1) tc::mma_commit_cg2(smem_u32(&acc_full[as])); // accumulator complete in BOTH CTAs
          }
          __syncwarp();
          if (++s == stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else {
    // ===== epilogue (both CTAs): rows m0 + 128 * rank + 32 * (w % 4) + lane of the pair tile, all BN columns =====
    const int q = warp & 3;
    bool al0, al1, alx;
    epilogue_alignment(p, al0, al1, alx);
    const uint32_t leader_empty0 = tc::mapa(smem_u32(&acc_empty[0]), 0), leader_empty1 = tc::mapa(smem_u32(&acc_empty[1]), 0);
    int local = 0;
    for (int tile = pair; tile < p.total_tiles; tile += npairs, ++local) {
      int m0, n0;
      tile_coords(tile, m0, n0);
      const int as = local & 1;
      tc::mbar_wait(smem_u32(&acc_full[as]), (uint32_t)(local >> 1) & 1u);
      tc::fence_after_sync();
      epilogue_tile(p, tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN), m0 + (int)rank * 128 + q * 32 + lane, n0, BN, al0, al1, alx);
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive_cluster(as ? leader_empty1 : leader_empty0);
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  tc::cluster_sync();            // nobody exits (or frees TMEM) while the peer may still touch its barriers / TMEM
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc_cg2(tmem_base, kTmemCols);
  }
}

template <int SA, int SB, int A_MN, int B_MN>
int launch(const GemmParams& p, int ctas, size_t smem, cudaStream_t st) {
  static std::atomic<unsigned long long> attr_done{0};
  if (int rc = ensure_dynamic_smem(gemm_tc_kernel<SA, SB, A_MN, B_MN>, (int)(kSmemBudget + 1024), attr_done)) return rc;
  PEA_CUDA(launch_ex(gemm_tc_kernel<SA, SB, A_MN, B_MN>, ctas, kThreads1, smem, st, p.pdl != 0, p));
  PEA_LAUNCH_CHECK();
  return PEA_OK;
}

template <int SA, int SB>
int dispatch_major(const GemmParams& p, int ctas, size_t smem, cudaStream_t st) {
  switch (p.a_mn * 2 + p.b_mn) {
    case 0: return launch<SA, SB, 0, 0>(p, ctas, smem, st);
    case 1: return launch<SA, SB, 0, 1>(p, ctas, smem, st);
    case 2: return launch<SA, SB, 1, 0>(p, ctas, smem, st);
    default: return launch<SA, SB, 1, 1>(p, ctas, smem, st);
  }
}

template <int A_MN, int B_MN>
int launch_pair(const GemmParams& p, int pairs, size_t smem, cudaStream_t st) {
  static std::atomic<unsigned long long> attr_done{0};
  if (int rc = ensure_dynamic_smem(gemm_tc2_kernel<A_MN, B_MN>, (int)(kSmemBudget + 1024), attr_done)) return rc;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = p.pdl ? 2 : 1;
  PEA_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc2_kernel<A_MN, B_MN>, p));
  PEA_LAUNCH_CHECK();
  return PEA_OK;
}

int dispatch_pair(const GemmParams& p, int pairs, size_t smem, cudaStream_t st) {
  switch (p.a_mn * 2 + p.b_mn) {
    case 0: return launch_pair<0, 0>(p, pairs, smem, st);
    case 1: return launch_pair<0, 1>(p, pairs, smem, st);
    case 2: return launch_pair<1, 0>(p, pairs, smem, st);
    default: return launch_pair<1, 1>(p, pairs, smem, st);
  }
}

int dispatch(const GemmParams& p, int ctas, size_t smem, cudaStream_t st) {
  switch (p.sa * 4 + p.sb) {
    case 1 * 4 + 1: return dispatch_major<1, 1>(p, ctas, smem, st);
    case 2 * 4 + 1: return dispatch_major<2, 1>(p, ctas, smem, st);
    case 1 * 4 + 2: return dispatch_major<1, 2>(p, ctas, smem, st);
    case 2 * 4 + 2: return dispatch_major<2, 2>(p, ctas, smem, st);
    case 3 * 4 + 3: return dispatch_major<3, 3>(p, ctas, smem, st);
  }
  set_error("gemm: plane combination (%d, %d) is not instantiated (supported: 1x1, 2x1, 1x2, 2x2, 3x3)", p.sa, p.sb);
  return PEA_EUNSUPPORTED;
}

}  // namespace

// ---- host: tensor-map encoding ------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

namespace {
struct TmapKey {
  const void* base; int rank; int fp16; uint64_t dims[3]; uint64_t strides[2]; uint32_t box[3]; int promo;
  bool operator==(const TmapKey& o) const { return memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(TmapKey) / 8; ++i) h = (h ^ w[i]) * 1099511628211ull;
    return (size_t)h;
  }
};
}  // namespace

int attn2_l2_promotion() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PEA_ATTN2_L2PROMO");
    v = e ? atoi(e) : 128;
  }
  return v;
}

// Descriptors are pure functions of (pointer, shape, strides, box): a training step re-launches the same ~250 kernels on
// the same buffers every iteration, so encodes (1-2 us of driver time each, 3-4 per launch) are cached process-wide.
int encode_tmap_16bit(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, bool fp16, int l2_promotion_bytes) {
  static EncodeTiledFn fn = nullptr;
  static std::mutex mu;
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.base = base; key.rank = rank; key.fp16 = fp16 ? 1 : 0; key.promo = l2_promotion_bytes;
  for (int i = 0; i < rank; ++i) { key.dims[i] = dims[i]; key.box[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) key.strides[i] = strides_bytes[i];
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return PEA_OK; }
  {
    // cuTensorMapEncodeTiled is a DRIVER call and needs a current context on the calling thread.  A thread whose first CUDA
    // work is one of our launches (an autograd worker starting a backward with a library GEMM) has none yet: the runtime
    // binds the primary context lazily, on its own first call.  Once per thread.
    static thread_local bool ctx_bound = false;
    if (!ctx_bound) { PEA_CUDA(cudaFree(nullptr)); ctx_bound = true; }
  }
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    PEA_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres));
    if (!sym || qres != cudaDriverEntryPointSuccess) {
      set_error("cuTensorMapEncodeTiled is not available from this driver");
      return PEA_ENODEV;
    }
    fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bdim[i] = box[i]; estr[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  const CUresult rc = fn(out, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank,
                         const_cast<void*>(base), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         l2_promotion_bytes >= 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                         : l2_promotion_bytes >= 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                         : l2_promotion_bytes >= 64  ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu x %llu, stride %llu, box %u x %u)", (int)rc, rank,
              (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 1),
              (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), box[0], rank > 1 ? box[1] : 1);
    return PEA_ECUDA;
  }
  if (cache.size() >= 8192) cache.clear();
  cache.emplace(key, *out);
  return PEA_OK;
}

}  // namespace pea

using namespace pea;

extern "C" int pea_gemm(const pea_gemm_desc* g, void* stream) {
  PEA_CHECK_ARG(g, "gemm: null descriptor");
  PEA_CHECK_ARG(g->M >= 1 && g->N >= 1 && g->K >= 1, "gemm: M, N, K must be >= 1 (got %d, %d, %d)", g->M, g->N, g->K);
  PEA_CHECK_ARG(g->in_dtype == PEA_BF16 || g->in_dtype == PEA_F16, "gemm: operands must be bf16 or fp16 planes");
  PEA_CHECK_ARG(g->nsplit_a >= 1 && g->nsplit_a <= 3 && g->nsplit_b >= 1 && g->nsplit_b <= 3, "gemm: nsplit_a/b must be 1..3");
  PEA_CHECK_ARG(g->lda % 8 == 0 && g->ldb % 8 == 0, "gemm: lda/ldb must be multiples of 8 elements (TMA 16-byte strides)");
  PEA_CHECK_ARG(g->lda >= (g->a_kmajor ? g->K : g->M) && g->ldb >= (g->b_kmajor ? g->K : g->N), "gemm: lda/ldb smaller than the row length");
  PEA_CHECK_ARG(g->d_out0 || g->out1_nsplit > 0, "gemm: no output requested");
  PEA_CHECK_ARG(g->out1_nsplit >= 0 && g->out1_nsplit <= 3, "gemm: out1_nsplit must be 0..3");
  PEA_CHECK_ARG(g->out1_nsplit == 0 || g->out1_dtype != PEA_F32, "gemm: out1 planes are 16-bit");
  PEA_CHECK_ARG(!g->d_rowadd || g->rowadd_group >= 1, "gemm: rowadd_group must be >= 1");

  GemmParams p{};
  p.M = g->M; p.N = g->N; p.K = g->K;
  p.sa = g->nsplit_a; p.sb = g->nsplit_b;
  p.a_mn = g->a_kmajor ? 0 : 1; p.b_mn = g->b_kmajor ? 0 : 1;
  const bool fp16 = g->in_dtype == PEA_F16;
  // Tile width: the multiple of 16 in [32, 256] that minimises  waves x (K blocks x time per stage + epilogue)  with
  //   time per stage = max(MMA: 2 bn clk per plane product, feed: stage bytes / ~48 B per clk of L2 -> shared memory per SM;
  //   calibrated on the 128 x 256 tile: 1 010 clk per K block at 4096^3).
  // Wide tiles halve the A re-reads and raise the tensor duty under the shared-memory bandwidth bound, but a one-wave
  // problem wants as many CTAs as fit in ONE wave: M = 1232, N = 2048 ran 160 tiles of 128 x 128 on 148 SMs -- twelve CTAs did
  // two -- and now runs 130 tiles of 128 x 160 (profiles/r2_gemm_timeline.txt).  PEA_GEMM_BN forces a width (A/B runs).
  const int m_tiles = (g->M + BM - 1) / BM;
  const int kb_all = (g->K + BK - 1) / BK;
  const int nsplit_k = g->split_k > 1 ? (g->split_k < kb_all ? g->split_k : kb_all) : 1;
  const int nprod = [&] { int n = 0; const int S = p.sa > p.sb ? p.sa : p.sb; for (int i = 0; i < p.sa; ++i) for (int j = 0; j < p.sb; ++j) n += (i + j < S); return n; }();
  int bn = 128;
  {
    static int forced = -2;
    if (forced == -2) { const char* e = getenv("PEA_GEMM_BN"); forced = e ? atoi(e) : -1; }
    double best = 1e30;
    const int sms = num_sms();
    for (int cand = 256; cand >= 32; cand -= 16) {
      const size_t stage = (size_t)p.sa * kTileBytes + (size_t)p.sb * (p.b_mn ? (size_t)((cand + 63) / 64) * 8192 : (size_t)cand * 128);
      if (kSmemBudget / stage < 2) continue;                       // needs at least a double buffer
      const long long tiles = (long long)m_tiles * ((g->N + cand - 1) / cand) * nsplit_k;
      const double waves = (double)((tiles + sms - 1) / sms);
      const double t_stage = fmax(2.0 * cand * nprod, (double)stage / 48.0);
      const double cost = waves * ((double)((kb_all + nsplit_k - 1) / nsplit_k) * t_stage + 12.0 * cand);
      if (cost < best * 0.999) { best = cost; bn = cand; }
    }
    if (forced >= 32 && forced <= 256 && forced % 16 == 0) bn = forced;
  }
  // CTA pairs (256 x 256 per two SMs) when there is a pair tile per SM pair, the row padding to 256 does not eat the gain
  // (measured: 1 329 vs 1 169 TF/s at 4096^3, x1.14; at M = 616 the 768-row padding loses: 895 vs 1 031 TF/s) AND the pair
  // tiles fill their last wave to >= 80 % (8192 x 1280 x 1280: 160 pair tiles on 74 pairs = 2.16 waves cost 37.8 us, the
  // searched 128 x 224 tiles 28 us).  PEA_GEMM_PAIR=0/1 forces the choice for A/B runs.
  bool pair = false;
  if (p.sa == 1 && p.sb == 1 && g->split_k <= 1) {   // the pair kernel is single-plane, no split-K
    const long long pm = (g->M + 255) / 256, pn = (g->N + 255) / 256;
    const int npairs = num_sms() / 2;
    const double eff_pair = (double)g->M / (256.0 * pm), eff_single = (double)g->M / (128.0 * m_tiles);
    const double wave_eff = (double)(pm * pn) / (double)(((pm * pn + npairs - 1) / npairs) * npairs);
    const double col_eff = (double)g->N / (256.0 * pn);     // N = 640 pads to 768 columns: 17 % of every pair tile wasted
    pair = pm * pn >= npairs && eff_pair * 1.12 > eff_single && wave_eff * col_eff >= 0.8;
    static int forced = -2;
    if (forced == -2) { const char* e = getenv("PEA_GEMM_PAIR"); forced = e ? atoi(e) : -1; }
    if (forced >= 0) pair = forced != 0;
    if (pair) bn = 256;
  }
  for (int i = 0; i < p.sa; ++i) {
    PEA_CHECK_ARG(g->d_A[i] && (reinterpret_cast<uintptr_t>(g->d_A[i]) & 15) == 0, "gemm: A plane %d null or not 16-byte aligned", i);
    const uint64_t dims[2] = {(uint64_t)(p.a_mn ? g->M : g->K), (uint64_t)(p.a_mn ? g->K : g->M)};
    const uint64_t str[1] = {(uint64_t)g->lda * 2};
    const uint32_t box[2] = {64u, p.a_mn ? 64u : 128u};
    int rc = encode_tmap_16bit(&p.tmA[i], g->d_A[i], 2, dims, str, box, fp16);
    if (rc != PEA_OK) return rc;
  }
  for (int j = 0; j < p.sb; ++j) {
    PEA_CHECK_ARG(g->d_B[j] && (reinterpret_cast<uintptr_t>(g->d_B[j]) & 15) == 0, "gemm: B plane %d null or not 16-byte aligned", j);
    const uint64_t dims[2] = {(uint64_t)(p.b_mn ? g->N : g->K), (uint64_t)(p.b_mn ? g->K : g->N)};
    const uint64_t str[1] = {(uint64_t)g->ldb * 2};
    const uint32_t box[2] = {64u, p.b_mn ? 64u : (pair ? 128u : (uint32_t)bn)};
    int rc = encode_tmap_16bit(&p.tmB[j], g->d_B[j], 2, dims, str, box, fp16);
    if (rc != PEA_OK) return rc;
  }
  p.kb_total = kb_all;
  int splits = g->split_k > 1 ? g->split_k : 1;
  if (splits > 1) {
    PEA_CHECK_ARG(g->d_out0 && g->out0_dtype == PEA_F32 && g->out0_accumulate && g->out1_nsplit == 0 && !g->d_bias && !g->d_aux && !g->d_rowadd,
                  "gemm: split_k needs an accumulating fp32 out0 and no bias/aux/rowadd/out1");
    if (splits > p.kb_total) splits = p.kb_total;
  }
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;  // no empty split
  p.m_tiles = pair ? (g->M + 255) / 256 : m_tiles;
  p.n_tiles = (g->N + bn - 1) / bn;
  p.splits = splits;
  p.total_tiles = p.m_tiles * p.n_tiles * splits;
  p.bn = bn;
  p.b_tile_bytes = p.b_mn ? (uint32_t)((bn + 63) / 64) * 8192u : (uint32_t)bn * 128u;
  const size_t stage_bytes = (size_t)p.sa * kTileBytes + (size_t)p.sb * p.b_tile_bytes;
  int stages = (int)(kSmemBudget / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  p.stages = stages;
  p.idesc = tc::make_idesc(fp16 ? 0 : 1, pair ? 256 : BM, bn, p.a_mn, p.b_mn);
  if (pair) {   // per-CTA stage = A 16 KiB + half of B 16 KiB
    stages = (int)(kSmemBudget / (2 * kTileBytes));
    if (stages > kMaxStages) stages = kMaxStages;
    p.stages = stages;
  }
  p.alpha = g->alpha;
  p.bias = g->d_bias;
  p.aux = g->d_aux; p.aux_dtype = g->aux_dtype; p.ldaux = g->ldaux;
  p.rowadd = g->d_rowadd; p.rowadd_group = g->rowadd_group; p.rowadd_scale = g->rowadd_scale;
  p.act = g->act;
  p.out0 = g->d_out0; p.out0_dtype = g->out0_dtype; p.ldo0 = g->ldo0;
  p.out0_mode = splits > 1 ? 2 : (g->out0_accumulate ? 1 : 0);
  for (int s = 0; s < g->out1_nsplit; ++s) {
    PEA_CHECK_ARG(g->d_out1[s], "gemm: out1 plane %d is null", s);
    p.out1[s] = g->d_out1[s];
  }
  p.out1_nsplit = g->out1_nsplit; p.out1_dtype = g->out1_dtype; p.ldo1 = g->ldo1;
  p.pdl = (g->flags & PEA_LAUNCH_PDL) ? 1 : 0;
  {
    static int disable = -1;   // PEA_GEMM_PLAIN_EPI=0: general epilogue everywhere (A/B runs)
    if (disable < 0) { const char* e = getenv("PEA_GEMM_PLAIN_EPI"); disable = (e && e[0] == '0') ? 1 : 0; }
    const uintptr_t o0 = reinterpret_cast<uintptr_t>(p.out0);
    p.epi_plain = !disable && p.out0 && p.out0_mode == 0 && p.out1_nsplit == 0 && !p.aux && !p.rowadd && g->N % 8 == 0 && p.ldo0 % 8 == 0 &&
                  (o0 & (p.out0_dtype == PEA_F32 ? 31 : 15)) == 0 && (!p.bias || (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
  }

  if (pair) {
    const int pairs = p.total_tiles < num_sms() / 2 ? p.total_tiles : num_sms() / 2;
    return dispatch_pair(p, pairs, (size_t)stages * 2 * kTileBytes + 1024, static_cast<cudaStream_t>(stream));
  }
  const size_t smem = (size_t)stages * stage_bytes + 1024;
  const int ctas = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  return dispatch(p, ctas, smem, static_cast<cudaStream_t>(stream));
}
