// attn2_tc.cu -- tcgen05 / TMEM / TMA forward of the fused cross-attention core for head_dim 64 (every SDXL attn2).
//
// Replaces F.scaled_dot_product_attention inside diffusers' AttnProcessor2_0 (reached from train_sdxl_zh.py:397,415 and
// tests/test_sdxl_zh.py:384-391) for the <= 80-key regime.  HBM-bound: Q streams in once, O streams out once; the
// projected context K|V of one (batch, head) -- 2 x 80 x 64 x 2 B = 20 KiB -- stays in shared memory while the CTA
// walks that head's query tiles.
//
// Persistent grid: one CTA per SM, each walking a CONTIGUOUS range of (batch, head, 128-row query tile) work items
// (query tile fastest), so the Q ring never drains and K|V is reloaded only when (batch, head) changes (double
// buffered, prefetched).  480 threads:
//   warp 0      TMA producer: K|V tiles (rows >= L zero-filled by TMA), Q tiles into a 4-deep ring
//   warp 14     S issuer:    S = Q K^T   tcgen05.mma SS, M128 N80 K64  -> TMEM S[t % 3], up to three tiles ahead
//   warp 1      TMEM allocator + P V issuer:
//                 O = P [V|1] tcgen05.mma TS, M128 N80 K80  -> TMEM O[t % 3]  (A = P straight from TMEM; the ones block
//                                                                          of B makes column 64 the row sum of P)
//   warps 2..13 three softmax/epilogue groups of four warps, tiles dealt round-robin; thread = query row = TMEM lane:
//               tcgen05.ld S, single-pass softmax (all keys are in the tile: no online rescaling and no cross-thread
//               reduction), unnormalised P -> 16-bit -> tcgen05.st over S's own columns, later tcgen05.ld O, scale by
//               1/rowsum, 128B-swizzled staging tile, TMA store (clips the N tail).
// Per tile a group is mostly waiting on round trips (TMEM ld/st, MMA commits: ~3500 clk of latency for ~1000 clk of issue;
// profiles/r1_attn2_fwd_timeline.txt), so three groups are in flight; S runs up to three tiles ahead of P V and the TMA
// engine up to four Q tiles ahead.  TMEM: 480 of 512 columns (S/P x3, O x3).
// Algorithmic bytes per tile: 128 x 64 x 2 B in + the same out; FLOPs 4 x 128 x 80 x 64.
#include "tc05.cuh"

namespace pea {
namespace {

constexpr int kQT = 128;                 // query rows per tile
constexpr int kD = 64;                   // head dim
constexpr int kLP = 80;                  // padded key count
#ifndef PEA_Q_STAGES
#define PEA_Q_STAGES 4
#endif
constexpr int kQStages = PEA_Q_STAGES;   // depth of the Q ring (16 KiB per slot; 7 is what shared memory allows)
constexpr int kGroups = 3;                    // softmax / epilogue groups (4 warps each), tiles dealt round-robin
constexpr int kSWarp = 2 + 4 * kGroups;       // the S = Q K^T issuer (warp 1 issues P V)
constexpr int kThreads = 64 + kGroups * 128 + 32;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kTmemS = 0, kTmemSStride = 80, kTmemO = 256, kTmemOStride = 80;   // 3 x 80 + 3 x 80 columns <= 512
constexpr int kNO = kD + 16;                 // O accumulator width: 64 value columns + the row-sum column (16: N granularity)
constexpr uint32_t kQBytes = kQT * kD * 2;      // 16 KiB
constexpr uint32_t kKVBytes = kLP * kD * 2;     // 10 KiB each (K, V)
constexpr uint32_t kKVBuf = 3 * kKVBytes;                      // K | V | ones (the row-sum column block of the B operand)
constexpr uint32_t kOffKV = 0;                                 // 2 buffers
constexpr uint32_t kOffQ = 2 * kKVBuf;                         // 4-deep ring
constexpr uint32_t kOffO = kOffQ + kQStages * kQBytes;         // one staging tile per softmax group
constexpr uint32_t kSmemBytes = kOffO + kGroups * kQBytes + 1024;   // one staging tile per softmax group + alignment slack

struct FwdParams {
  CUtensorMap tmQ, tmKV, tmO;
  int B, N, L, H, C;
  int tiles_per_bh, total_tiles, tiles_per_cta;
  float scale_log2;
  uint32_t idesc_s, idesc_o;
};

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// exp2 of two values at once on the MUFU's packed 16-bit path: the result is exactly the 16-bit P pair the tensor core
// consumes (P is rounded to 16 bits anyway), at half the MUFU issue slots of two fp32 ex2 + a convert.
template <typename T> __device__ __forceinline__ uint32_t ex2_pair(float x0, float x1);
template <> __device__ __forceinline__ uint32_t ex2_pair<__nv_bfloat16>(float x0, float x1) {
  uint32_t r;
  const uint32_t x = pack2<__nv_bfloat16>(x0, x1);
  asm("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(r) : "r"(x));
  return r;
}
template <> __device__ __forceinline__ uint32_t ex2_pair<__half>(float x0, float x1) {
  uint32_t r;
  const uint32_t x = pack2<__half>(x0, x1);
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(r) : "r"(x));
  return r;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// Optional per-tile event timeline of CTA 0 (build with -DPEA_TRACE, read with tools/trace_fwd.py).  No-op otherwise.
#ifdef PEA_TRACE
__device__ unsigned long long g_trace_fwd[16 * 64 * 16];   // [role][tile][event]
#define TRACE(role, tile, ev) do { if (blockIdx.x == 0 && (tile) < 64 && lane == 0) g_trace_fwd[((role) * 64 + (tile)) * 16 + (ev)] = clock64(); } while (0)
#else
#define TRACE(role, tile, ev) do { } while (0)
#endif

template <typename T>
__global__ void __launch_bounds__(kThreads, 1) attn2_fwd_tc_kernel(const __grid_constant__ FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t kv_full[2], kv_empty[2], q_full[kQStages], q_empty[kQStages];
  __shared__ __align__(8) uint64_t s_full[kGroups], p_ready[kGroups], o_full[kGroups], tmem_free[kGroups];
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  if (warp == 0) TRACE(0, 0, 15);   // kernel entry
  TRACE(15, warp, 0);               // per-warp entry

  // ---- this CTA's contiguous range of work items; item = bh * tiles_per_bh + tile ----------------------------------
  const int item0 = blockIdx.x * p.tiles_per_cta;
  const int ntiles = min(p.tiles_per_cta, p.total_tiles - item0);
  const int tpb = p.tiles_per_bh;
  // Work items are consecutive, so every role walks (batch, head, tile) with counters; the only integer divisions of the
  // kernel are these.  (The MMA warp is a single latency-bound instruction stream that shares its scheduler with three
  // softmax warps: four divisions per tile in it were a third of the tile period.)
  const int bh0 = fast_div(item0, tpb), rem0 = item0 - bh0 * tpb;
  const int b0 = fast_div(bh0, p.H), h0 = bh0 - b0 * p.H;

#ifdef PEA_EXP_PARINIT
  if (warp == 0) {
    // one barrier per lane: kv_full[2] kv_empty[2] | s_full, p_ready, o_full, tmem_free [kGroups each] | q_full, q_empty [kQStages each]
    if (lane < 3) tc::prefetch_tmap(lane == 0 ? (const void*)&p.tmQ : lane == 1 ? (const void*)&p.tmKV : (const void*)&p.tmO);
    if (lane < 2) tc::mbar_init(smem_u32(&kv_full[lane]), 1);
    else if (lane < 4) tc::mbar_init(smem_u32(&kv_empty[lane - 2]), 1);
    else if (lane < 4 + kGroups) tc::mbar_init(smem_u32(&s_full[lane - 4]), 1);
    else if (lane < 4 + 2 * kGroups) tc::mbar_init(smem_u32(&p_ready[lane - 4 - kGroups]), 4);
    else if (lane < 4 + 3 * kGroups) tc::mbar_init(smem_u32(&o_full[lane - 4 - 2 * kGroups]), 1);
    else if (lane < 4 + 4 * kGroups) tc::mbar_init(smem_u32(&tmem_free[lane - 4 - 3 * kGroups]), 4);
    else if (lane < 4 + 4 * kGroups + kQStages) tc::mbar_init(smem_u32(&q_full[lane - 4 - 4 * kGroups]), 1);
    else if (lane < 4 + 4 * kGroups + 2 * kQStages) tc::mbar_init(smem_u32(&q_empty[lane - 4 - 4 * kGroups - kQStages]), 1);
    tc::fence_mbar_init();
  }
#else
  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&p.tmQ); tc::prefetch_tmap(&p.tmKV); tc::prefetch_tmap(&p.tmO);
#pragma unroll
    for (int i = 0; i < 2; ++i) { tc::mbar_init(smem_u32(&kv_full[i]), 1); tc::mbar_init(smem_u32(&kv_empty[i]), 1); }
#pragma unroll
    for (int i = 0; i < kGroups; ++i) {
      tc::mbar_init(smem_u32(&s_full[i]), 1); tc::mbar_init(smem_u32(&p_ready[i]), 4);
      tc::mbar_init(smem_u32(&o_full[i]), 1); tc::mbar_init(smem_u32(&tmem_free[i]), 4);
    }
#pragma unroll
    for (int i = 0; i < kQStages; ++i) { tc::mbar_init(smem_u32(&q_full[i]), 1); tc::mbar_init(smem_u32(&q_empty[i]), 1); }
    tc::fence_mbar_init();
  }
#endif
  if (warp == 1) tc::tmem_alloc(smem_u32(&tmem_slot), kTmemCols);
  if (warp >= 2) {
    // B operand of O = P V is [V | 1]: columns 64..79 of the accumulator then hold the row sum of the ROUNDED P the tensor
    // core actually multiplies, so the softmax threads never add it up.  All-ones is swizzle invariant: plain fill.
    const uint32_t one2 = pack2<T>(1.0f, 1.0f);
    for (uint32_t i = threadIdx.x - 64; i < 2 * kKVBytes / 16; i += kThreads - 64) {
      const uint32_t a = smem + kOffKV + (i / (kKVBytes / 16)) * kKVBuf + 2 * kKVBytes + (i % (kKVBytes / 16)) * 16;
      asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(a), "r"(one2) : "memory");
    }
    tc::fence_proxy_async_smem();
  }
  TRACE(15, warp, 1);               // per-warp: own prologue work done
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  TRACE(15, warp, 2);               // per-warp: past the prologue barrier
  if (warp == 0) TRACE(0, 0, 14);   // prologue done (barriers, TMEM)
  pdl_wait();                 // everything above overlapped the preceding kernel's tail (PEA_LAUNCH_PDL); Q / K|V are its output
  pdl_launch_dependents();

  if (warp == 0) {
    // ===== TMA producer =====
    int kv_n = 0;
    int b = b0, h = h0, tile = rem0;
    for (int t = 0; t < ntiles; ++t) {
      if (t == 0 || tile == 0) {   // new (batch, head): next K|V buffer (freed by the last P V that read it two loads ago)
        const int kb = kv_n & 1;
        tc::mbar_wait(smem_u32(&kv_empty[kb]), ((uint32_t)(kv_n >> 1) & 1u) ^ 1u);
        if (tc::elect_one()) {
          const uint32_t bk = smem_u32(&kv_full[kb]);
          tc::mbar_expect_tx(bk, 2 * kKVBytes);
          tc::tma_load_3d(smem + kOffKV + kb * kKVBuf, &p.tmKV, bk, h * kD, 0, b);
          tc::tma_load_3d(smem + kOffKV + kb * kKVBuf + kKVBytes, &p.tmKV, bk, p.C + h * kD, 0, b);
        }
        __syncwarp();
        if (t == 0) TRACE(0, 0, 13);   // first K|V loads issued
        ++kv_n;
      }
      const int s = t % kQStages;
      if (t >= kQStages) tc::mbar_wait(smem_u32(&q_empty[s]), ((uint32_t)(t / kQStages) & 1u) ^ 1u);   // the first ring pass is free:
      TRACE(0, t, 0);                                                                                  // no ~100-clk try_wait per start-up load
      if (tc::elect_one()) {
        const uint32_t fb = smem_u32(&q_full[s]);
        tc::mbar_expect_tx(fb, kQBytes);
        tc::tma_load_3d(smem + kOffQ + s * kQBytes, &p.tmQ, fb, h * kD, tile * kQT, b);
      }
      __syncwarp();
      if (++tile == tpb) { tile = 0; if (++h == p.H) { h = 0; ++b; } }
    }
  } else if (warp == 1) {
    // ===== P V issuer =====
    // MN-major tile (V as B of P V): [key][d], 16 keys per K step = 2048 B.
    constexpr uint64_t kDescMN = ((uint64_t)(kKVBytes >> 4) << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
    int o_pos = rem0, o_kvn = 0;    // tile-in-head position and K|V bookkeeping of the P V stream
    int o_j = 0; uint32_t o_ph = 0; // t % kGroups and (t / kGroups) & 1
    for (int t = 0; t < ntiles; ++t) {
      const int j = o_j;
      const uint32_t ph = o_ph;
      if (++o_j == kGroups) { o_j = 0; o_ph ^= 1u; }
      if (t == 0 || o_pos == 0) ++o_kvn;
      const int kb = (o_kvn - 1) & 1;
      const bool last_of_bh = (t + 1 < ntiles) && (o_pos == tpb - 1);
      if (++o_pos == tpb) o_pos = 0;
      TRACE(1, t, 2);
      tc::mbar_wait(smem_u32(&p_ready[j]), ph);
      TRACE(1, t, 3);
      if (t >= kGroups) tc::mbar_wait(smem_u32(&tmem_free[j]), ph ^ 1u);  // O[j] of tile t-3 read out
      tc::fence_after_sync();
      if (tc::elect_one()) {
        const uint32_t va = (smem + kOffKV + kb * kKVBuf + kKVBytes) >> 4;
        const uint32_t tp = tmem + kTmemS + j * kTmemSStride, to = tmem + kTmemO + j * kTmemOStride;
#pragma unroll
        for (int k = 0; k < kLP / 16; ++k)  // A = P from TMEM (8 packed columns per K=16 step), B = V
          tc::mma_f16_ts(to, tp + k * 8, kDescMN | (uint64_t)(va + k * 128), p.idesc_o, k > 0);
        tc::mma_commit(smem_u32(&o_full[j]));
        if (last_of_bh) tc::mma_commit(smem_u32(&kv_empty[kb]));   // K|V buffer reusable once these MMAs have read it
      }
      __syncwarp();
    }
  } else if (warp == kSWarp) {
    // ===== S issuer.  S and P V are issued by two warps: one latency-bound instruction stream issuing both (and sharing
    // its scheduler with three softmax warps) paced the whole pipeline -- 18.4 -> 16.4 us at N = 1024, C = 1280.  Ordering
    // with the P V warp is explicit: S(t) reuses S/P[t % 3], free once P V(t-3) has completed (o_full); a K|V buffer is
    // released by the last P V of its head (kv_empty) and S of the head after next simply waits for the reload, so any
    // number of tiles per head works without special cases. =====
    // K-major tiles (Q, K): 128-byte rows, 8-row groups 1024 B apart; K=16 step = 32 B inside the swizzle atom.
    constexpr uint64_t kDescK = ((uint64_t)1 << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
    int s_pos = rem0, s_kvn = 0, j = 0;
    uint32_t jph = 0;
    for (int t = 0; t < ntiles; ++t) {
      if (t == 0 || s_pos == 0) {
        tc::mbar_wait(smem_u32(&kv_full[s_kvn & 1]), (uint32_t)(s_kvn >> 1) & 1u);
        if (t == 0) TRACE(1, 0, 5);   // first K|V landed
        ++s_kvn;
      }
      if (++s_pos == tpb) s_pos = 0;
      const int kb = (s_kvn - 1) & 1, qs = t % kQStages;
      TRACE(1, t, 0);
      tc::mbar_wait(smem_u32(&q_full[qs]), (uint32_t)(t / kQStages) & 1u);
      TRACE(1, t, 1);
      if (t >= kGroups) tc::mbar_wait(smem_u32(&o_full[j]), jph ^ 1u);   // P V(t-3) done with P in S/P[j]
      tc::fence_after_sync();
      if (tc::elect_one()) {
        const uint32_t qa = (smem + kOffQ + qs * kQBytes) >> 4, ka = (smem + kOffKV + kb * kKVBuf) >> 4;
#pragma unroll
        for (int k = 0; k < kD / 16; ++k)
          tc::mma_f16_ss(tmem + kTmemS + j * kTmemSStride, kDescK | (uint64_t)(qa + k * 2), kDescK | (uint64_t)(ka + k * 2), p.idesc_s, k > 0);
        tc::mma_commit(smem_u32(&s_full[j]));
        tc::mma_commit(smem_u32(&q_empty[qs]));
      }
      __syncwarp();
      if (++j == kGroups) { j = 0; jph ^= 1u; }
    }
  } else {
    // ===== softmax + epilogue groups: warp w owns TMEM lanes [32*(w%4), +32); thread = query row =====
    const int g = (warp - 2) >> 2;                       // group g: warps 2+4g .. 5+4g, tiles g, g+3, g+6, ...
    const int q = warp & 3;
    const int row = q * 32 + lane;                       // row inside the tile
    const uint32_t t_s = tmem + kTmemS + g * kTmemSStride + ((uint32_t)(q * 32) << 16);
    const uint32_t t_o = tmem + kTmemO + g * kTmemOStride + ((uint32_t)(q * 32) << 16);
    const bool issuer = threadIdx.x == 64 + g * 128;
    const int L = p.L;
    const float sl2 = p.scale_log2;
    int b = b0, h = h0, tile = rem0 + g;   // (batch, head, tile) of this group's current item, advanced kGroups at a time
    uint32_t ph = 0;
    for (int t = g; t < ntiles; t += kGroups, ph ^= 1u, tile += kGroups) {
      while (tile >= tpb) { tile -= tpb; if (++h == p.H) { h = 0; ++b; } }
      TRACE(2 + g, t, 0);
      tc::mbar_wait(smem_u32(&s_full[g]), ph);
      TRACE(2 + g, t, 1);
      tc::fence_after_sync();
      uint32_t sv[kLP];
      {
        uint32_t(&c0)[16] = *reinterpret_cast<uint32_t(*)[16]>(&sv[0]);
        uint32_t(&c1)[16] = *reinterpret_cast<uint32_t(*)[16]>(&sv[16]);
        uint32_t(&c2)[16] = *reinterpret_cast<uint32_t(*)[16]>(&sv[32]);
        uint32_t(&c3)[16] = *reinterpret_cast<uint32_t(*)[16]>(&sv[48]);
        uint32_t(&c4)[16] = *reinterpret_cast<uint32_t(*)[16]>(&sv[64]);
        tc::tmem_ld16(t_s + 0, c0); tc::tmem_ld16(t_s + 16, c1); tc::tmem_ld16(t_s + 32, c2);
        tc::tmem_ld16(t_s + 48, c3); tc::tmem_ld16(t_s + 64, c4);
        tc::tmem_ld_wait();
      }
      // columns [0, Lfull) need no masking (Lfull = L rounded down to 16: warp-uniform branches below)
      // four independent chains: a single softmax warp per scheduler cannot hide a serial 80-long dependency chain
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int c = 0; c < kLP / 16; ++c) {
        if ((c + 1) * 16 <= L) {
#pragma unroll
          for (int j = 0; j < 16; ++j) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(sv[c * 16 + j]));
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) m4[j & 3] = fmaxf(m4[j & 3], c * 16 + j < L ? __uint_as_float(sv[c * 16 + j]) : -INFINITY);
        }
      }
      const float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      const float off = m * sl2;
#pragma unroll
      for (int c = 0; c < kLP / 16; ++c) {
        uint32_t pk[8];
        if ((c + 1) * 16 <= L) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            pk[j] = ex2_pair<T>(fmaf(__uint_as_float(sv[c * 16 + 2 * j]), sl2, -off), fmaf(__uint_as_float(sv[c * 16 + 2 * j + 1]), sl2, -off));
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int j0 = c * 16 + 2 * j;   // masked keys: exp2(-inf) = 0
            pk[j] = ex2_pair<T>(j0 < L ? fmaf(__uint_as_float(sv[j0]), sl2, -off) : -INFINITY,
                                j0 + 1 < L ? fmaf(__uint_as_float(sv[j0 + 1]), sl2, -off) : -INFINITY);
          }
        }
        tmem_st8(t_s + c * 8, pk);   // P (16-bit pairs) over S's own columns: this thread's row only
      }
      tc::tmem_st_wait();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(smem_u32(&p_ready[g]));
      TRACE(2 + g, t, 2);

      // staging tile free again? (this group's previous TMA store must have finished READING it)
      const uint32_t o_stage = smem + kOffO + (uint32_t)g * kQBytes;
      if (t >= kGroups) {
        if (issuer) tc::tma_store_wait_read<0>();
        named_bar_sync(1 + g, 128);
      }
      TRACE(2 + g, t, 3);
      tc::mbar_wait(smem_u32(&o_full[g]), ph);
      TRACE(2 + g, t, 4);
      tc::fence_after_sync();
      {
        uint32_t ov[64];
        uint32_t(&o0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&ov[0]);
        uint32_t(&o1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&ov[32]);
        uint32_t rs[8];                       // column 64 = sum_j P_j (columns 65..79 repeat it)
        tc::tmem_ld32(t_o, o0);
        tc::tmem_ld32(t_o + 32, o1);
        tmem_ld8(t_o + 64, rs);
        tc::tmem_ld_wait();
        const float inv = 1.0f / __uint_as_float(rs[0]);
#pragma unroll
        for (int chunk = 0; chunk < 8; ++chunk) {   // 16-byte chunk of the 128-byte row; TMA SWIZZLE_128B: chunk ^= row % 8
          uint4 u;
          u.x = pack2<T>(__uint_as_float(ov[chunk * 8 + 0]) * inv, __uint_as_float(ov[chunk * 8 + 1]) * inv);
          u.y = pack2<T>(__uint_as_float(ov[chunk * 8 + 2]) * inv, __uint_as_float(ov[chunk * 8 + 3]) * inv);
          u.z = pack2<T>(__uint_as_float(ov[chunk * 8 + 4]) * inv, __uint_as_float(ov[chunk * 8 + 5]) * inv);
          u.w = pack2<T>(__uint_as_float(ov[chunk * 8 + 6]) * inv, __uint_as_float(ov[chunk * 8 + 7]) * inv);
          const uint32_t addr = o_stage + (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4);
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(smem_u32(&tmem_free[g]));   // S / P / O[g] may be overwritten by tile t + 2
      tc::fence_proxy_async_smem();                              // generic-proxy smem writes -> visible to the TMA engine
      named_bar_sync(1 + g, 128);
      if (issuer) {
        tc::tma_store_3d(&p.tmO, o_stage, h * kD, tile * kQT, b);
        tc::tma_store_commit();
      }
      TRACE(2 + g, t, 5);
    }
#ifdef PEA_EXP_WAITREAD
    if (issuer) tc::tma_store_wait_read<0>();   // shared memory must outlive the store's READ; the global writes complete by grid end
#else
    if (issuer) tc::tma_store_wait<0>();
#endif
    TRACE(2 + g, 63, 6);   // group done (stores complete for the issuer's lane 0)
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) TRACE(0, 0, 12);   // all roles done
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem, kTmemCols);
  }
}

template <typename T>
int launch_fwd_tc(const FwdParams& p, int units, cudaStream_t st, bool pdl) {
  static std::atomic<unsigned long long> attr_done{0};
  if (int rc = ensure_dynamic_smem(attn2_fwd_tc_kernel<T>, (int)kSmemBytes, attr_done)) return rc;
  PEA_CUDA(launch_ex(attn2_fwd_tc_kernel<T>, units, kThreads, kSmemBytes, st, pdl, p));
  PEA_LAUNCH_CHECK();
  return PEA_OK;
}

}  // namespace

#ifdef PEA_TRACE
extern "C" int pea_debug_fwd_trace(unsigned long long* host, int n) {
  return (int)cudaMemcpyFromSymbol(host, g_trace_fwd, sizeof(unsigned long long) * n);
}
#endif

// Called by pea_attn2_fwd (attn2_mma.cu) when head_dim == 64 and the layout fits TMA (16-byte strides).
int attn2_fwd_tc(const pea_attn2_desc* a, cudaStream_t st) {
  FwdParams p{};
  const bool fp16 = a->dtype == PEA_F16;
  const int C = a->H * a->d;
  p.B = a->B; p.N = a->N; p.L = a->L; p.H = a->H; p.C = C;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  {
    const uint64_t dims[3] = {(uint64_t)C, (uint64_t)a->N, (uint64_t)a->B};
    const uint64_t str[2] = {(uint64_t)a->ld_q * 2, (uint64_t)a->ld_q * 2 * (uint64_t)a->N};
    const uint32_t box[3] = {64, 128, 1};
    int rc = encode_tmap_16bit(&p.tmQ, a->d_q, 3, dims, str, box, fp16, attn2_l2_promotion());
    if (rc != PEA_OK) return rc;
    const uint64_t ostr[2] = {(uint64_t)a->ld_o * 2, (uint64_t)a->ld_o * 2 * (uint64_t)a->N};
    rc = encode_tmap_16bit(&p.tmO, a->d_o, 3, dims, ostr, box, fp16, attn2_l2_promotion());
    if (rc != PEA_OK) return rc;
  }
  {
    // K and V share one map: V = K's base + (d_v - d_k) elements along the row (the K|V interleaved layout)
    const long long voff = (static_cast<const char*>(a->d_v) - static_cast<const char*>(a->d_k)) / 2;
    const uint64_t dims[3] = {(uint64_t)(voff + C), (uint64_t)a->L, (uint64_t)a->B};
    const uint64_t str[2] = {(uint64_t)a->ld_kv * 2, (uint64_t)a->bs_kv * 2};
    const uint32_t box[3] = {64, (uint32_t)kLP, 1};
    int rc = encode_tmap_16bit(&p.tmKV, a->d_k, 3, dims, str, box, fp16, attn2_l2_promotion());
    if (rc != PEA_OK) return rc;
    p.C = (int)voff;  // column offset of V inside the shared map
  }
  p.tiles_per_bh = (a->N + kQT - 1) / kQT;
  p.total_tiles = a->B * a->H * p.tiles_per_bh;
  // one persistent CTA per SM, each with a contiguous range of work items
  const int slots = num_sms();
  p.tiles_per_cta = (p.total_tiles + slots - 1) / slots;
  const int units = (p.total_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
  p.idesc_s = tc::make_idesc(fp16 ? 0 : 1, 128, kLP, 0, 0);
  p.idesc_o = tc::make_idesc(fp16 ? 0 : 1, 128, kNO, 0, 1);
  const bool pdl = (a->flags & PEA_LAUNCH_PDL) != 0;
  return fp16 ? launch_fwd_tc<__half>(p, units, st, pdl) : launch_fwd_tc<__nv_bfloat16>(p, units, st, pdl);
}

bool attn2_fwd_tc_eligible(const pea_attn2_desc* a) {
  if (a->d != kD || a->L > kLP) return false;
  const long long voff = (static_cast<const char*>(a->d_v) - static_cast<const char*>(a->d_k)) / 2;
  const int C = a->H * a->d;
  // K|V interleaved in one [B, L, *] tensor with V to the right of K inside the row, 16-byte aligned everything
  return voff >= C && voff + C <= a->ld_kv && a->ld_q % 8 == 0 && a->ld_o % 8 == 0 && a->ld_kv % 8 == 0 && a->bs_kv % 8 == 0 &&
         (reinterpret_cast<uintptr_t>(a->d_q) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->d_k) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(a->d_o) & 15) == 0 && voff % 8 == 0;
}

}  // namespace pea
