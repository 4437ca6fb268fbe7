// attn2_bwd_tc.cu -- tcgen05 / TMEM / TMA backward of the fused cross-attention core for head_dim 64.
//
// Gradient of softmax(q k^T scale) v as autograd forms it for diffusers' AttnProcessor2_0 under train_sdxl_zh.py:397
// (student UNet; DeepSpeed backward) -- dQ, and dK / dV which the caller turns into dContext.  Because all <= 80 keys
// sit in one tile, S and P are RECOMPUTED from Q, K and delta = rowsum(P o dP) is local to a query row, so neither O
// nor a log-sum-exp is read: HBM traffic is read Q, read dO, write dQ (+ the tiny K, V, dK, dV).
//
// One persistent CTA per SM walks a contiguous range of (batch, head, 128-row query tile) items.  512 threads in four
// warpgroups whose register budgets are re-balanced with setmaxnreg (softmax 176, dQ epilogue 96, the rest 56):
//   WG0  warp 0   TMA producer: K|V (double buffered, on a (batch, head) change), Q and dO tiles (4-deep ring)
//        warp 1   TMEM allocator + MMA issuer A, per tile t:  dQ(t) = dS K   TS N64 K80  (A = 16-bit dS in TMEM over dP[g])
//                                                              then S(t+2) = Q K^T, dP(t+2) = dO V^T   SS N80 K64 -> S[g], dP[g]
//        warp 2   dQ TMA-store issuer
//        warp 3   MMA issuer B:  dV^T += dO^T P, dK^T += Q^T dS   SS M64 N80 K128 (A = the dO / Q tile read MN-major, B = P / dS);
//                 released by issuer A right after dQ(t), so its 16 MMAs interleave with S / dP(t+2) in the tensor pipe
//                 instead of sitting in front of them (one issuing warp) or behind them (which starves the next P / dS stores)
//   WG1  warps 4-7   softmax group 0 (even tiles)     thread = query row = TMEM lane:
//   WG2  warps 8-11  softmax group 1 (odd tiles)      P = softmax(S), delta, dS = P o (dP - delta) scale; P, dS -> shared
//                 memory and dS -> TMEM; on a (batch, head) change the group drains dK^T / dV^T and adds them to the global
//                 fp32 accumulators (red.global.add.f32, 16 consecutive floats per half warp and key)
//   WG3  warps 12-15 dQ epilogue: TMEM -> 16-bit -> 128B-swizzled staging tile (then warp 2's TMA store), so a softmax group
//                 goes straight from dS(t) to S / dP(t+2)
// Round-2 restructuring (profiles/r2_attn2_bwd_timeline.txt): in round 1 a group's cycle was softmax(t) 2 850 clk -> wait dQ
// 720 -> dQ epilogue 950 -> wait S / dP(t+2) 720, the last because S / dP(t+2) queued behind dV^T / dK^T(t) in the in-order
// tensor pipe.  Now the dQ epilogue has its own warpgroup, S / dP(t+2) and dV^T / dK^T(t) are issued by two warps at the same
// time, and the Q / dO ring is 4 deep (~4 200 cycles of HBM latency under load need >= 3 tiles in flight per SM) -- the 4th
// slot bought by storing P / dS in the un-swizzled ("interleaved") MN-major layout: 8-key x 128-row columns of 16-byte units,
// 20 KiB each and perfectly coalesced row-per-thread stores, instead of two 64-key 128B-swizzle slabs (32 KiB).
// dK / dV are formed TRANSPOSED (rows = head dim, columns = keys): an M64 MMA fetches 2 KB of A + 2.5 KB of B per K=16 step;
// an M64 accumulator occupies lanes 0-15 of each 32-lane quarter, so dK^T interleaves into lanes 16-31 of dV^T's 80 columns.
// They accumulate in TMEM across the CTA's consecutive tiles of one (batch, head).
// TMEM (512 columns): S[2] | dP[2] (160 each pair), dQ 64, dV^T / dK^T 80 (lane-interleaved).
// Algorithmic bytes per tile: 2 x 16 KiB in (Q, dO) + 16 KiB out (dQ); FLOPs 10 x 128 x 80 x 64.
#include "tc05.cuh"

namespace pea {
namespace {

constexpr int kQT = 128, kD = 64, kLP = 80;
constexpr int kRing = 4;
constexpr int kThreads = 512;
constexpr int kRegsSoftmax = 176, kRegsEpilogue = 96, kRegsOther = 56;   // 128 * 56 + 256 * 176 + 128 * 96 <= 65 536
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kTS = 0, kTdP = 80, kTPair = 160, kTdQ = 320, kTdKV = 384;
constexpr uint32_t kLaneDK = 16u << 16;        // dK^T sits 16 lanes below dV^T in every quarter
constexpr uint32_t kTile = kQT * kD * 2;        // 16 KiB: one Q / dO / dQ tile
constexpr uint32_t kKV = kLP * kD * 2;          // 10 KiB: K or V
constexpr uint32_t kPCol = kQT * 16;            // 2 KiB: one 8-key column of P / dS ([128 q][8 keys], 16 B per row)
constexpr uint32_t kPBytes = (kLP / 8) * kPCol; // 20 KiB
constexpr uint32_t kOffKV = 0;                              // 2 x (K | V)
constexpr uint32_t kOffRing = 2 * 2 * kKV;                  // 4 x (Q | dO)
constexpr uint32_t kOffP = kOffRing + kRing * 2 * kTile;
constexpr uint32_t kOffdS = kOffP + kPBytes;
constexpr uint32_t kOffStage = kOffdS + kPBytes;            // dQ staging
constexpr uint32_t kSmemBytes = kOffStage + kTile + 1024;   // 225 KiB
static_assert(kSmemBytes + 512 <= 227 * 1024, "shared memory budget");

struct BwdParams {
  CUtensorMap tmQ, tmdO, tmKV, tmdQ;
  int B, N, L, H, voff;
  int tiles_per_bh, total_tiles, tiles_per_cta;
  float scale_log2, scale;
  uint32_t idesc_s, idesc_dq, idesc_dkv;
  float* dk; float* dv; long long ld_dkv, bs_dkv;
};

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void st_shared16(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

// Optional per-tile event timeline of CTA 0 (build with -DPEA_TRACE, read with tools/trace_bwd.py).  Expands to nothing in
// the product build.
#ifdef PEA_TRACE
__device__ unsigned long long g_trace[16 * 64 * 16];   // [role][tile][event]
#define TRACE(role, tile, ev) do { if (blockIdx.x == 0 && (tile) < 64 && lane == 0) g_trace[((role) * 64 + (tile)) * 16 + (ev)] = clock64(); } while (0)
#else
#define TRACE(role, tile, ev) do { } while (0)
#endif

template <typename T>
__global__ void __launch_bounds__(kThreads, 1) attn2_bwd_tc_kernel(const __grid_constant__ BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t kv_full[2], kv_empty[2], ring_full[kRing], ring_empty[kRing];
  __shared__ __align__(8) uint64_t sdp_full[2], ds_ready[2], pds_free, dq_full, dq_free, dkv_full, dkv_free, dq_staged, stage_free, go_dkv;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int item0 = blockIdx.x * p.tiles_per_cta;
  const int ntiles = min(p.tiles_per_cta, p.total_tiles - item0);
  const int tpb = p.tiles_per_bh;
  // Work items are consecutive: every role walks (batch, head, tile) with counters; the only divisions of the kernel are these
  const int bh0 = fast_div(item0, tpb), rem0 = item0 - bh0 * tpb;
  const int b0 = fast_div(bh0, p.H), h0 = bh0 - b0 * p.H;
  if (warp == 0) TRACE(0, 63, 15);   // kernel entry

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&p.tmQ); tc::prefetch_tmap(&p.tmdO); tc::prefetch_tmap(&p.tmKV); tc::prefetch_tmap(&p.tmdQ);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(smem_u32(&kv_full[i]), 1); tc::mbar_init(smem_u32(&kv_empty[i]), 1);
      tc::mbar_init(smem_u32(&sdp_full[i]), 1); tc::mbar_init(smem_u32(&ds_ready[i]), 4);
    }
#pragma unroll
    for (int i = 0; i < kRing; ++i) { tc::mbar_init(smem_u32(&ring_full[i]), 1); tc::mbar_init(smem_u32(&ring_empty[i]), 1); }
    tc::mbar_init(smem_u32(&pds_free), 1);
    tc::mbar_init(smem_u32(&dq_full), 1);
    tc::mbar_init(smem_u32(&dq_free), 4);
    tc::mbar_init(smem_u32(&dkv_full), 1);
    tc::mbar_init(smem_u32(&dkv_free), 4);
    tc::mbar_init(smem_u32(&dq_staged), 4);
    tc::mbar_init(smem_u32(&stage_free), 1);
    tc::mbar_init(smem_u32(&go_dkv), 1);
    tc::fence_mbar_init();
  }
  if (warp == 1) tc::tmem_alloc(smem_u32(&tmem_slot), kTmemCols);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  pdl_wait();                 // see attn2_tc.cu: the prologue overlapped the preceding kernel's tail under PEA_LAUNCH_PDL
  pdl_launch_dependents();

  // K-major tiles [rows][64 x 16-bit] (Q, dO as A; K, V as B of S / dP): K=16 step = 32 B inside the 128B swizzle atom.
  // MN-major 128B-swizzle tiles: K as B of dQ, Q / dO as A of dK^T / dV^T ([k rows][64 contiguous]): K=16 step = 16 rows = 2048 B.
  // MN-major un-swizzled ("interleaved") P / dS as B of dV^T / dK^T: core matrix = 8 q-rows x 8 keys = 128 contiguous bytes;
  //   next 8 keys (SBO) 2 KiB further, next 8 q-rows (LBO) 128 B further; K=16 step = 2 core matrices = 256 B.
  constexpr uint64_t kDescK = ((uint64_t)1 << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
  constexpr uint64_t kDescMN = ((uint64_t)1 << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
  constexpr uint64_t kDescPdS = ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(kPCol >> 4) << 32) | (1ull << 46);

  if (warp < 4) {
    reg_dec<kRegsOther>();
    if (warp == 0) {
      // ===== TMA producer =====
      int kv_n = 0, s = 0;
      uint32_t ring_ph = 0;
      int b = b0, h = h0, tile = rem0;
      for (int t = 0; t < ntiles; ++t) {
        if (t == 0 || tile == 0) {   // first tile of a (batch, head) in this CTA's range
          const int kb = kv_n & 1;
          tc::mbar_wait(smem_u32(&kv_empty[kb]), ((uint32_t)(kv_n >> 1) & 1u) ^ 1u);
          if (tc::elect_one()) {
            const uint32_t bk = smem_u32(&kv_full[kb]);
            tc::mbar_expect_tx(bk, 2 * kKV);
            tc::tma_load_3d(smem + kOffKV + kb * 2 * kKV, &p.tmKV, bk, h * kD, 0, b);
            tc::tma_load_3d(smem + kOffKV + kb * 2 * kKV + kKV, &p.tmKV, bk, p.voff + h * kD, 0, b);
          }
          __syncwarp();
          ++kv_n;
        }
        if (t >= kRing) tc::mbar_wait(smem_u32(&ring_empty[s]), ring_ph ^ 1u);   // the first ring pass is free
        TRACE(0, t, 0);
        if (tc::elect_one()) {
          const uint32_t fb = smem_u32(&ring_full[s]);
          tc::mbar_expect_tx(fb, 2 * kTile);
          tc::tma_load_3d(smem + kOffRing + s * 2 * kTile, &p.tmQ, fb, h * kD, tile * kQT, b);
          tc::tma_load_3d(smem + kOffRing + s * 2 * kTile + kTile, &p.tmdO, fb, h * kD, tile * kQT, b);
        }
        __syncwarp();
        if (++s == kRing) { s = 0; ring_ph ^= 1u; }
        if (++tile == tpb) { tile = 0; if (++h == p.H) { h = 0; ++b; } }
      }
    } else if (warp == 1) {
      // ===== MMA issuer A: dQ(t), then S / dP(t+2) =====
      int s_pos = rem0, s_kvn = 0;   // tile-in-head position / K|V bookkeeping of the S / dP stream (two tiles ahead)
      int m_pos = rem0, m_kvn = 0;   // ... and of the dQ stream
      int s_rs = 0;                  // ring slot (t % kRing) of the S / dP stream
      uint32_t s_rph = 0;            // (t / kRing) & 1
      auto issue_sdp = [&](int t) {
        if (t == 0 || s_pos == 0) {
          tc::mbar_wait(smem_u32(&kv_full[s_kvn & 1]), (uint32_t)(s_kvn >> 1) & 1u);
          ++s_kvn;
        }
        if (++s_pos == tpb) s_pos = 0;
        const int kb = (s_kvn - 1) & 1, rs = s_rs, g = t & 1;
        const uint32_t rph = s_rph;
        if (++s_rs == kRing) { s_rs = 0; s_rph ^= 1u; }
        TRACE(1, t, 0);
        tc::mbar_wait(smem_u32(&ring_full[rs]), rph);
        TRACE(1, t, 1);
        tc::fence_after_sync();
        if (tc::elect_one()) {
          const uint32_t qa = (smem + kOffRing + rs * 2 * kTile) >> 4, doa = qa + (kTile >> 4);
          const uint32_t ka = (smem + kOffKV + kb * 2 * kKV) >> 4, va = ka + (kKV >> 4);
          const uint32_t ts = tmem + g * kTPair + kTS, tdp = tmem + g * kTPair + kTdP;
#pragma unroll
          for (int k = 0; k < kD / 16; ++k) {
            tc::mma_f16_ss(ts, kDescK | (uint64_t)(qa + k * 2), kDescK | (uint64_t)(ka + k * 2), p.idesc_s, k > 0);
            tc::mma_f16_ss(tdp, kDescK | (uint64_t)(doa + k * 2), kDescK | (uint64_t)(va + k * 2), p.idesc_s, k > 0);
          }
          tc::mma_commit(smem_u32(&sdp_full[g]));
        }
        __syncwarp();
      };
      if (ntiles > 0) issue_sdp(0);
      if (ntiles > 1) issue_sdp(1);
      for (int t = 0; t < ntiles; ++t) {
        const int g = t & 1;
        if (t == 0 || m_pos == 0) ++m_kvn;
        const int kb = (m_kvn - 1) & 1;
        const bool last_of_bh = (t + 1 < ntiles) && (m_pos == tpb - 1);
        if (++m_pos == tpb) m_pos = 0;
        TRACE(1, t, 2);
        tc::mbar_wait(smem_u32(&ds_ready[g]), (uint32_t)(t >> 1) & 1u);
        TRACE(1, t, 3);
        if (t > 0) tc::mbar_wait(smem_u32(&dq_free), (uint32_t)(t - 1) & 1u);   // dQ TMEM drained by the epilogue warpgroup
        TRACE(1, t, 4);
        tc::fence_after_sync();
        if (tc::elect_one()) {
          const uint32_t ka = (smem + kOffKV + kb * 2 * kKV) >> 4;
          const uint32_t tds = tmem + g * kTPair + kTdP;
          // dQ = dS K   (A: 16-bit dS in TMEM, 8 columns per K=16 step; B: K tile read MN-major)
#pragma unroll
          for (int k = 0; k < kLP / 16; ++k)
            tc::mma_f16_ts(tmem + kTdQ, tds + k * 8, kDescMN | (uint64_t)(ka + k * 128), p.idesc_dq, k > 0);
          tc::mma_commit(smem_u32(&dq_full));
          // S, dP and dQ are the only readers of K|V, and dQ of a head's last tile is the last of them (its S / dP completed
          // before its softmax did): buffer reusable.  (dV^T / dK^T read only the Q / dO ring and P / dS.)
          if (last_of_bh) tc::mma_commit(smem_u32(&kv_empty[kb]));
        }
        __syncwarp();
        // Issuer B queues dV^T / dK^T(t) while this warp issues S / dP(t+2): the two streams interleave in the tensor pipe.
        // (S / dP(t+2) strictly first was measured slower, 261 vs 232 us at N = 16384: dV^T / dK^T(t) then complete so late that
        // the next tile's P / dS stores -- one shared-memory copy -- wait for them; profiles/r2_attn2_bwd_timeline.txt.)
        if (lane == 0) tc::mbar_arrive(smem_u32(&go_dkv));
        if (t + 2 < ntiles) issue_sdp(t + 2);     // same thread as dQ(t): dP(t+2) overwrites dP[g] after dQ(t) has consumed dS
      }
    } else if (warp == 3) {
      // ===== MMA issuer B: dV^T += dO^T P ; dK^T += Q^T dS (reduction over the 128 query rows; the two chains interleaved) =====
      int m_pos = rem0, m_rs = 0, flushes = 0;
      for (int t = 0; t < ntiles; ++t) {
        const int g = t & 1, rs = m_rs;
        if (++m_rs == kRing) m_rs = 0;
        const bool first_of_bh = t == 0 || m_pos == 0;
        const bool last_tile = t + 1 == ntiles;
        const bool last_of_bh = !last_tile && (m_pos == tpb - 1);
        if (++m_pos == tpb) m_pos = 0;
        tc::mbar_wait(smem_u32(&ds_ready[g]), (uint32_t)(t >> 1) & 1u);        // P / dS of tile t are in shared memory
        tc::mbar_wait(smem_u32(&go_dkv), (uint32_t)t & 1u);
        if (first_of_bh && flushes > 0) tc::mbar_wait(smem_u32(&dkv_free), (uint32_t)(flushes - 1) & 1u);  // dK/dV TMEM drained
        tc::fence_after_sync();
        if (tc::elect_one()) {
          const uint32_t qa = (smem + kOffRing + rs * 2 * kTile) >> 4, doa = qa + (kTile >> 4);
          const uint32_t pa = (smem + kOffP) >> 4, dsa = (smem + kOffdS) >> 4;
#pragma unroll
          for (int k = 0; k < kQT / 16; ++k) {
            tc::mma_f16_ss(tmem + kTdKV, kDescMN | (uint64_t)(doa + k * 128), kDescPdS | (uint64_t)(pa + k * 16), p.idesc_dkv,
                           (k > 0) ? 1u : (uint32_t)!first_of_bh);
            tc::mma_f16_ss(tmem + kTdKV + kLaneDK, kDescMN | (uint64_t)(qa + k * 128), kDescPdS | (uint64_t)(dsa + k * 16), p.idesc_dkv,
                           (k > 0) ? 1u : (uint32_t)!first_of_bh);
          }
          tc::mma_commit(smem_u32(&pds_free));            // P / dS shared-memory tiles reusable
          tc::mma_commit(smem_u32(&ring_empty[rs]));      // Q / dO slot reusable
          if (last_of_bh || last_tile) tc::mma_commit(smem_u32(&dkv_full));
        }
        __syncwarp();
        if (last_of_bh || last_tile) ++flushes;
      }
    } else {
      // ===== warp 2: dQ TMA-store issuer =====
      int b = b0, h = h0, tile = rem0;
      for (int t = 0; t < ntiles; ++t) {
        tc::mbar_wait(smem_u32(&dq_staged), (uint32_t)t & 1u);
        if (tc::elect_one()) {
          tc::tma_store_3d(&p.tmdQ, smem + kOffStage, h * kD, tile * kQT, b);
          tc::tma_store_commit();
          tc::tma_store_wait_read<0>();
          tc::mbar_arrive(smem_u32(&stage_free));
          // (the global writes of the last store complete by grid end; shared memory only has to outlive the READ)
        }
        __syncwarp();
        if (++tile == tpb) { tile = 0; if (++h == p.H) { h = 0; ++b; } }
      }
    }
  } else if (warp < 12) {
    // ===== softmax groups =====
    reg_inc<kRegsSoftmax>();
    const int g = (warp - 4) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t t_s = tmem + g * kTPair + kTS + lane_off, t_dp = tmem + g * kTPair + kTdP + lane_off;
    const uint32_t row_off = (uint32_t)row * 16u;       // this row's 16-byte unit inside every 8-key column of P / dS
    const int L = p.L;
    const float sl2 = p.scale_log2, scale = p.scale;
    int flushes = 0;   // counted identically in both groups and MMA issuer B
    int nb = b0, nh = h0, pos = rem0;   // (batch, head, tile-in-head) of tile t, advanced once per tile
    for (int t = 0; t < ntiles; ++t) {
      const int b = nb, h = nh;
      const bool last_tile = t + 1 == ntiles;
      const bool flush = last_tile || (pos == tpb - 1);
      if (++pos == tpb) { pos = 0; if (++nh == p.H) { nh = 0; ++nb; } }
      if ((t & 1) != g) {            // the other group's tile: only keep the flush counter in step
        if (flush) ++flushes;
        continue;
      }
      TRACE(2 + g, t, 0);
      tc::mbar_wait(smem_u32(&sdp_full[g]), (uint32_t)(t >> 1) & 1u);
      TRACE(2 + g, t, 1);
      tc::fence_after_sync();

      // Columns >= L hold S = 0 (K rows beyond L are zero-filled), never NaN.
      // ---- pass A: S -> e = exp2((s - max) * scale * log2 e), row sum ---------------------------------------------
      float e[kLP];
      {
        uint32_t sv[kLP];
        uint32_t(&c0)[16] = *reinterpret_cast<uint32_t(*)[16]>(&sv[0]);
        uint32_t(&c1)[16] = *reinterpret_cast<uint32_t(*)[16]>(&sv[16]);
        uint32_t(&c2)[16] = *reinterpret_cast<uint32_t(*)[16]>(&sv[32]);
        uint32_t(&c3)[16] = *reinterpret_cast<uint32_t(*)[16]>(&sv[48]);
        uint32_t(&c4)[16] = *reinterpret_cast<uint32_t(*)[16]>(&sv[64]);
        tc::tmem_ld16(t_s + 0, c0); tc::tmem_ld16(t_s + 16, c1); tc::tmem_ld16(t_s + 32, c2);
        tc::tmem_ld16(t_s + 48, c3); tc::tmem_ld16(t_s + 64, c4);
        tc::tmem_ld_wait();
        // keys >= L -> -inf once (see attn2_tc.cu): max / exp below then exist in one unmasked copy; exp2(-inf) = 0
#pragma unroll
        for (int c = 0; c < kLP / 16; ++c) {
          if ((c + 1) * 16 > L) {
#pragma unroll
            for (int j = 0; j < 16; ++j) sv[c * 16 + j] = c * 16 + j < L ? sv[c * 16 + j] : 0xff800000u;
          }
        }
        // four independent chains: a single softmax warp per scheduler cannot hide a serial 80-long dependency chain
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < kLP; ++j) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(sv[j]));
        const float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        const float off = m * sl2;
#pragma unroll
        for (int j = 0; j < kLP; ++j) e[j] = ex2(fmaf(__uint_as_float(sv[j]), sl2, -off));
      }
      float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < kLP; ++j) s4[j & 3] += e[j];
      const float inv = 1.0f / ((s4[0] + s4[1]) + (s4[2] + s4[3]));
      TRACE(2 + g, t, 8);

      // ---- pass B: delta = sum_j P_j dP_j.  TMEM loads are software-pipelined: chunk c+1 is in flight while chunk c
      //      is consumed (tcgen05.wait::ld right after every load would expose its full latency ten times per tile).
      float d4[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t dva[16], dvb[16];
      tc::tmem_ld16(t_dp, dva);
      tc::tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < kLP / 16; ++c) {
        uint32_t(&cur)[16] = (c & 1) ? dvb : dva;
        uint32_t(&nxt)[16] = (c & 1) ? dva : dvb;
        tc::tmem_ld16(t_dp + ((c + 1 < kLP / 16) ? (c + 1) * 16 : 0), nxt);   // last iteration prefetches pass C's chunk 0
#pragma unroll
        for (int j = 0; j < 16; ++j) d4[j & 3] = fmaf(e[c * 16 + j], __uint_as_float(cur[j]), d4[j & 3]);
        tc::tmem_ld_wait();
      }
      const float delta = ((d4[0] + d4[1]) + (d4[2] + d4[3])) * inv;
      TRACE(2 + g, t, 9);

      // ---- pass C: dS = P o (dP - delta) * scale, P: 16-bit pairs in registers; dS -> TMEM over dP -------------------
      const float nds = -delta * scale;   // dS = P * (dP * scale - delta * scale): one FFMA + one FMUL per element
      uint32_t pp[kLP / 2], dd[kLP / 2];
#pragma unroll
      for (int c = 0; c < kLP / 16; ++c) {
        // kLP / 16 = 5 is odd, so pass B's last prefetch (chunk 0) landed in dvb: chunk c lives in (c even ? dvb : dva)
        uint32_t(&cur)[16] = (c & 1) ? dva : dvb;
        uint32_t(&nxt)[16] = (c & 1) ? dvb : dva;
        if (c + 1 < kLP / 16) tc::tmem_ld16(t_dp + (c + 1) * 16, nxt);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float p0 = e[c * 16 + 2 * j] * inv, p1 = e[c * 16 + 2 * j + 1] * inv;
          pp[c * 8 + j] = pack2<T>(p0, p1);
          dd[c * 8 + j] = pack2<T>(p0 * fmaf(__uint_as_float(cur[2 * j]), scale, nds), p1 * fmaf(__uint_as_float(cur[2 * j + 1]), scale, nds));
        }
        if (c + 1 < kLP / 16) tc::tmem_ld_wait();   // chunk c+1 (columns >= 16(c+1)) is in registers before columns [8c, 8c+8) are overwritten
        uint32_t(&d8)[8] = *reinterpret_cast<uint32_t(*)[8]>(&dd[c * 8]);
        tmem_st8(t_dp + c * 8, d8);
      }

      // ---- P / dS shared-memory tiles free?  (dV^T / dK^T of the previous tile have read them) -- only the stores wait
      TRACE(2 + g, t, 2);
      if (t > 0) tc::mbar_wait(smem_u32(&pds_free), (uint32_t)(t - 1) & 1u);
      TRACE(2 + g, t, 3);
#pragma unroll
      for (int chunk = 0; chunk < kLP / 8; ++chunk) {          // 8 keys = 16 bytes; un-swizzled MN-major: [8-key column][q row]
        const uint32_t o = (uint32_t)chunk * kPCol + row_off;  // a warp writes 512 contiguous bytes per store
        st_shared16(smem + kOffP + o, pp[chunk * 4 + 0], pp[chunk * 4 + 1], pp[chunk * 4 + 2], pp[chunk * 4 + 3]);
        st_shared16(smem + kOffdS + o, dd[chunk * 4 + 0], dd[chunk * 4 + 1], dd[chunk * 4 + 2], dd[chunk * 4 + 3]);
      }
      tc::tmem_st_wait();
      tc::fence_before_sync();
      tc::fence_proxy_async_smem();     // P / dS written through the generic proxy -> visible to tcgen05.mma
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(smem_u32(&ds_ready[g]));
      TRACE(2 + g, t, 4);

      // ---- (batch, head) boundary: drain dV^T | dK^T (lane = head-dim element) into the global fp32 accumulators -----
      if (flush) {
        tc::mbar_wait(smem_u32(&dkv_full), (uint32_t)flushes & 1u);
        tc::fence_after_sync();
        uint32_t acc[kLP];
        {
          uint32_t(&a0)[16] = *reinterpret_cast<uint32_t(*)[16]>(&acc[0]);
          uint32_t(&a1)[16] = *reinterpret_cast<uint32_t(*)[16]>(&acc[16]);
          uint32_t(&a2)[16] = *reinterpret_cast<uint32_t(*)[16]>(&acc[32]);
          uint32_t(&a3)[16] = *reinterpret_cast<uint32_t(*)[16]>(&acc[48]);
          uint32_t(&a4)[16] = *reinterpret_cast<uint32_t(*)[16]>(&acc[64]);
          const uint32_t ta = tmem + kTdKV + lane_off;
          tc::tmem_ld16(ta + 0, a0); tc::tmem_ld16(ta + 16, a1); tc::tmem_ld16(ta + 32, a2);
          tc::tmem_ld16(ta + 48, a3); tc::tmem_ld16(ta + 64, a4);
          tc::tmem_ld_wait();
        }
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0 && !last_tile) tc::mbar_arrive(smem_u32(&dkv_free));   // TMEM dK / dV may be overwritten by the next head
        {
          // quarter q holds rows 16q .. 16q+15 of dV^T in lanes 0-15 and of dK^T in lanes 16-31
          float* ptr = (lane < 16 ? p.dv : p.dk) + (long long)b * p.bs_dkv + h * kD + q * 16 + (lane & 15);
          const long long ld = p.ld_dkv;
#pragma unroll
          for (int key = 0; key < kLP; ++key) {     // running pointer: 3-4 instructions per key instead of a 64-bit multiply each
            if (key < L) atomicAdd(ptr, __uint_as_float(acc[key]));
            ptr += ld;
          }
        }
        ++flushes;
      }
    }
    TRACE(2 + g, 63, 14);   // group done (incl. dK/dV flush)
  } else {
    // ===== warps 12-15: dQ epilogue: TMEM -> 16-bit -> swizzled staging tile -> (warp 2) TMA store =====
    reg_dec<kRegsEpilogue>();
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    for (int t = 0; t < ntiles; ++t) {
      tc::mbar_wait(smem_u32(&dq_full), (uint32_t)t & 1u);
      tc::fence_after_sync();
      uint32_t ov[64];
      uint32_t(&o0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&ov[0]);
      uint32_t(&o1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&ov[32]);
      tc::tmem_ld32(tmem + kTdQ + lane_off, o0);
      tc::tmem_ld32(tmem + kTdQ + lane_off + 32, o1);
      tc::tmem_ld_wait();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(smem_u32(&dq_free));      // dQ TMEM may be overwritten by tile t + 1
      if (t > 0) tc::mbar_wait(smem_u32(&stage_free), (uint32_t)(t - 1) & 1u);   // previous store has read the staging tile
#pragma unroll
      for (int chunk = 0; chunk < 8; ++chunk) {
        const uint32_t addr = smem + kOffStage + (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4);
        st_shared16(addr, pack2<T>(__uint_as_float(ov[chunk * 8 + 0]), __uint_as_float(ov[chunk * 8 + 1])),
                    pack2<T>(__uint_as_float(ov[chunk * 8 + 2]), __uint_as_float(ov[chunk * 8 + 3])),
                    pack2<T>(__uint_as_float(ov[chunk * 8 + 4]), __uint_as_float(ov[chunk * 8 + 5])),
                    pack2<T>(__uint_as_float(ov[chunk * 8 + 6]), __uint_as_float(ov[chunk * 8 + 7])));
      }
      tc::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(smem_u32(&dq_staged));
      TRACE(4, t, 0);
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) TRACE(0, 63, 14);   // all roles done
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem, kTmemCols);
  }
}

template <typename T>
int launch_bwd_tc(const BwdParams& p, int ctas, cudaStream_t st, bool pdl) {
  static std::atomic<unsigned long long> attr_done{0};
  if (int rc = ensure_dynamic_smem(attn2_bwd_tc_kernel<T>, (int)kSmemBytes, attr_done)) return rc;
  PEA_CUDA(launch_ex(attn2_bwd_tc_kernel<T>, ctas, kThreads, kSmemBytes, st, pdl, p));
  PEA_LAUNCH_CHECK();
  return PEA_OK;
}

}  // namespace

#ifdef PEA_TRACE
extern "C" int pea_debug_bwd_trace(unsigned long long* host, int n) {
  return (int)cudaMemcpyFromSymbol(host, g_trace, sizeof(unsigned long long) * n);
}
#endif

bool attn2_bwd_tc_eligible(const pea_attn2_desc* a) {
  if (a->d != kD || a->L > kLP) return false;
  const long long voff = (static_cast<const char*>(a->d_v) - static_cast<const char*>(a->d_k)) / 2;
  const int C = a->H * a->d;
  auto al16 = [](const void* x) { return (reinterpret_cast<uintptr_t>(x) & 15) == 0; };
  return voff >= C && voff + C <= a->ld_kv && voff % 8 == 0 && a->ld_q % 8 == 0 && a->ld_do % 8 == 0 && a->ld_dq % 8 == 0 &&
         a->ld_kv % 8 == 0 && a->bs_kv % 8 == 0 && al16(a->d_q) && al16(a->d_k) && al16(a->d_do) && al16(a->d_dq);
  // dK / dV are fp32 accumulators updated with scalar red.global.add.f32: no layout constraint beyond the ABI's
}

int attn2_bwd_tc(const pea_attn2_desc* a, cudaStream_t st) {
  BwdParams p{};
  const bool fp16 = a->dtype == PEA_F16;
  const int C = a->H * a->d;
  p.B = a->B; p.N = a->N; p.L = a->L; p.H = a->H;
  p.scale = a->scale;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  const uint64_t dims[3] = {(uint64_t)C, (uint64_t)a->N, (uint64_t)a->B};
  const uint32_t box[3] = {64, 128, 1};
  struct { CUtensorMap* m; const void* ptr; long long ld; } maps[3] = {{&p.tmQ, a->d_q, a->ld_q}, {&p.tmdO, a->d_do, a->ld_do}, {&p.tmdQ, a->d_dq, a->ld_dq}};
  for (auto& m : maps) {
    const uint64_t str[2] = {(uint64_t)m.ld * 2, (uint64_t)m.ld * 2 * (uint64_t)a->N};
    int rc = encode_tmap_16bit(m.m, m.ptr, 3, dims, str, box, fp16, attn2_l2_promotion());
    if (rc != PEA_OK) return rc;
  }
  {
    const long long voff = (static_cast<const char*>(a->d_v) - static_cast<const char*>(a->d_k)) / 2;
    const uint64_t kdims[3] = {(uint64_t)(voff + C), (uint64_t)a->L, (uint64_t)a->B};
    const uint64_t str[2] = {(uint64_t)a->ld_kv * 2, (uint64_t)a->bs_kv * 2};
    const uint32_t kbox[3] = {64, (uint32_t)kLP, 1};
    int rc = encode_tmap_16bit(&p.tmKV, a->d_k, 3, kdims, str, kbox, fp16, attn2_l2_promotion());
    if (rc != PEA_OK) return rc;
    p.voff = (int)voff;
  }
  p.tiles_per_bh = (a->N + kQT - 1) / kQT;
  p.total_tiles = a->B * a->H * p.tiles_per_bh;
  const int slots = num_sms();
  p.tiles_per_cta = (p.total_tiles + slots - 1) / slots;
  const int ctas = (p.total_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
  const int fmt = fp16 ? 0 : 1;
  p.idesc_s = tc::make_idesc(fmt, 128, kLP, 0, 0);
  p.idesc_dq = tc::make_idesc(fmt, 128, kD, 0, 1);
  p.idesc_dkv = tc::make_idesc(fmt, 64, kLP, 1, 1);
  p.dk = a->d_dk; p.dv = a->d_dv; p.ld_dkv = a->ld_dkv; p.bs_dkv = a->bs_dkv;
  const bool pdl = (a->flags & PEA_LAUNCH_PDL) != 0;
  return fp16 ? launch_bwd_tc<__half>(p, ctas, st, pdl) : launch_bwd_tc<__nv_bfloat16>(p, ctas, st, pdl);
}

}  // namespace pea
