// Multi-head self-attention BACKWARD on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), for the short CLIP
// sequences (vision L = 50+P, text L ≤ 77; head dim 64).  Given Q, K, V (the taped in-proj output) and dO:
//   S = Q·Kᵀ/8 (+causal mask), P = softmax(S), dP = dO·Vᵀ, dS = P∘(dP − rowsum(P∘dP)),
//   dQ = dS·K/8, dK = dSᵀ·Q/8, dV = Pᵀ·dO          per (sample, head).
// Reference: autograd through nn.MultiheadAttention inside clip.model.ResidualAttentionBlock (reached from
// models/clip_encoders.py:75-84,186) w.r.t. the prompt rows (models/prompts_models.py:52) — the data gradient of
// the frozen attention.
//
// One work item = one head of G samples (G = 2 while L ≤ 64, else 1), packed as in attn_tc.cu: sample j of the item
// occupies rows [64·j, 64·j + L) of four 128-row shared-memory tiles Q, K, V, dO (128-byte rows, 128B swizzle,
// written by TMA boxes of exactly L rows; the rows in between stay zero).  Five tcgen05.mma chains per item:
//   S  = Q·Kᵀ    128×128×64   (both operands K-major)            → TMEM columns [0,128)
//   dP = dO·Vᵀ   128×128×64                                        → TMEM [128,256)
//   dV = Pᵀ·dO   128×64×128   A = P read MN-major (transposed), B = dO read MN-major     → TMEM [384,448)
//   dK = dSᵀ·Q   128×64×128   A = dS MN-major, B = Q MN-major                             → TMEM [320,384)
//   dQ = dS·K    128×64×128   A = dS K-major,  B = K MN-major                             → TMEM [256,320)
// P and dS are fp16 128×128 tiles in shared memory whose off-diagonal 64×64 quadrants (query of one sample × key of
// the other) are zero, which keeps the samples apart in all three products.
// Warp-specialised persistent CTAs: warp 0 TMA producer (2- or 3-stage ring of Q|K|V|dO), warp 1 MMA issuer, warp 2 TMEM
// allocator; warpgroup 1 (thread = query row) turns S, dP into P, dS; warpgroup 2 (thread = output row) drains
// dQ, dK, dV → fp16 → staging (the dead P/dS tiles) → TMA stores of L rows per sample.  The two accumulator groups
// (S, dP | dQ, dK, dV) are disjoint, so item i+1's S/dP and its soft-max overlap item i's second-phase MMAs and stores.
// Algorithmic HBM traffic: 14·D bytes per token (read q, k, v, dO; write dq, dk, dv).
#include <stdlib.h>

#include "common.cuh"
#include "ctx.h"

using namespace gb;

// Optional stall accounting (build with -DGB_ATTN_STALLS; read back with gb_debug_attn_stalls): clocks per role spent
// waiting on each barrier, summed over CTAs.  [0] MMA: operands, [1] MMA: S/dP free, [2] MMA: P/dS ready, [3] MMA: outputs
// free, [4] MMA total; [5] soft-max wg: S/dP complete, [6] its P/dS region free, [7] its total; [8] output wg: outputs
// complete, [9] its store-read wait, [10] its total; [11] producer: free stage; [12] items.
#ifdef GB_ATTN_STALLS
__device__ unsigned long long g_attn_stalls[16];
#define AB_T(t) const long long t = clock64()
#define AB_ADD(i, t) atomicAdd(&g_attn_stalls[i], (unsigned long long)(clock64() - t))
#define AB_INC(i) atomicAdd(&g_attn_stalls[i], 1ull)
#else
#define AB_T(t)
#define AB_ADD(i, t)
#define AB_INC(i)
#endif

namespace {

constexpr int kAbTile = 128 * 128;             // one 128-row × 64-half tile, bytes
constexpr int kAbPBytes = 2 * kAbTile;         // P (and dS): two 64-key atoms of 128 rows
// Input ring geometry: four full 128-row tiles per stage, two stages (a third does not fit beside the P / dS tiles and the
// output staging; tried with short tiles at one sample per item, where it fits — no gain: once the issue order is
// dynamic the producer waits for a free stage most of the time, the soft-max warpgroup's arithmetic is what an item costs).
template <int G, int NB>
struct AbGeom {
  static constexpr int kRows = 128;
  static constexpr int kTile = kRows * 128;
  static constexpr int kStageBytes = 4 * kTile;      // Q | K | V | dO
  static constexpr int kStages = 2;
  static constexpr int kKeys = 128;                  // N of the S / dP products
  static constexpr int kSmem = kStages * kStageBytes + 2 * kAbPBytes + 2 * kAbTile + 1024 + 256;
};
constexpr int kAbThreads = 128 + 2 * 128;
constexpr int kAbStagingTiles = 2;             // fp16 output staging, ping-pong
static_assert(kAbStagingTiles == 2, "AbGeom::kSmem counts two staging tiles");
constexpr uint32_t kColS = 0, kColdP = 128, kColdQ = 256, kColdK = 320, kColdV = 384;

struct AttnBwdParams {
  int B, L, H, D;
  int causal;
  int n_groups;   // ceil(B / G)
  int n_items;    // n_groups · H
};

// Operand read MN-major (rows of the tile = K dimension of the product, 64 contiguous M/N elements per 128-byte row,
// 128B swizzle): groups of 8 K-rows are 1024 B apart (SBO); 64-element blocks along M/N are `lbo` bytes apart.
__device__ __forceinline__ uint64_t ab_desc_mn(uint32_t smem_addr, uint32_t lbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__device__ __forceinline__ void ab_wg_barrier(int id) {
  asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory");
}

// 16 consecutive accumulator columns into 16 registers (no wait: several are issued back to back — a tcgen05.ld that
// has to wait for running MMAs takes 1-2 k clocks to deliver, so that latency is paid once per item, not per piece)
__device__ __forceinline__ void ab_tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// G samples per item; NB = ceil(L / 16) blocks of 16 key columns a query row looks at (and of 16 rows per sample that
// the products over queries / keys have to visit)
template <int G, int NB>
__global__ void __launch_bounds__(kAbThreads, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                   const __grid_constant__ CUtensorMap tmDQKV, const AttnBwdParams p) {
  constexpr int kCols = 16 * NB;              // key columns a query row looks at
  static_assert(G == 2 ? NB <= 4 : (NB > 4 && NB <= 6), "G = 2 packs samples of up to 64 rows; G = 1 up to 96");
  using Geo = AbGeom<G, NB>;
  constexpr int kAbStages = Geo::kStages;
  constexpr int kAbStageBytes = Geo::kStageBytes;
  constexpr int kInTile = Geo::kTile;         // bytes of one input tile (Q, K, V or dO) of a stage
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_in = smem;                                        // [stage][Q|K|V|dO][128][128 B]
  uint8_t* smem_p = smem_in + kAbStages * kAbStageBytes;          // [atom][128][128 B]
  uint8_t* smem_ds = smem_p + kAbPBytes;                          // [atom][128][128 B]
  uint8_t* smem_stg = smem_ds + kAbPBytes;                        // [kAbStagingTiles][128][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stg + kAbStagingTiles * kAbTile);
  uint64_t* full_bar = bars;                    // [kAbStages]
  uint64_t* empty_bar = bars + kAbStages;       // [kAbStages]
  uint64_t* sdp_full = bars + 2 * kAbStages;    // S and dP complete
  uint64_t* sdp_free = sdp_full + 1;            // S and dP read out of TMEM
  uint64_t* pds_ready = sdp_free + 1;           // P and dS in shared memory
  uint64_t* out_full = pds_ready + 1;           // dQ, dK, dV complete (and with them P, dS dead)
  uint64_t* out_free = out_full + 1;            // dQ, dK, dV read out of TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(out_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmDQKV);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kAbStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(sdp_full, 1);
    mbar_init(sdp_free, 128);
    mbar_init(pds_ready, 128);
    mbar_init(out_full, 1);
    mbar_init(out_free, 128);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  // rows a TMA box never writes (≥ L in each 64-row half) must read as zero
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const int n16 = (kAbStages * kAbStageBytes + 2 * kAbPBytes) / 16;
    for (int i = threadIdx.x; i < n16; i += kAbThreads) z[i] = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int L = p.L, H = p.H, D = p.D;

  // register budgets per warpgroup (384 threads: 128·(56 + 224 + 224) = 64 512 ≤ 65 536)
  if (warp < 4) reg_dealloc<56>();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int it = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
        const int s = it % kAbStages;
        const uint32_t ph = (it / kAbStages) & 1;
        const int grp = item / H, h = item % H;
        const int n_s = min(G, p.B - grp * G);
        { AB_T(t0); mbar_wait(&empty_bar[s], ph ^ 1); AB_ADD(11, t0); }
        mbar_expect_tx(&full_bar[s], 4u * n_s * L * 128u);
        uint8_t* st = smem_in + s * kAbStageBytes;
        for (int j = 0; j < n_s; ++j) {
          const int row0 = (grp * G + j) * L;
          for (int m = 0; m < 3; ++m)
            tma_load_2d(st + m * kInTile + j * 8192, &tmQKV, &full_bar[s], m * D + h * 64, row0);
          tma_load_2d(st + 3 * kInTile + j * 8192, &tmDO, &full_bar[s], h * 64, row0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_f16(128, Geo::kKeys);                       // A, B K-major
      constexpr uint32_t idesc_t = umma_idesc_f16(128, 64) | (1u << 15) | (1u << 16);     // A, B MN-major
      constexpr uint32_t idesc_q = umma_idesc_f16(128, 64) | (1u << 16);                  // A K-major, B MN-major
      const uint32_t pa = smem_u32(smem_p), da = smem_u32(smem_ds);
      const int n_mine = blockIdx.x < p.n_items ? (p.n_items - 1 - blockIdx.x) / (int)gridDim.x + 1 : 0;
      AB_T(t_mma);
      // first phase of item `it`: S = Q·Kᵀ and dP = dO·Vᵀ into their TMEM columns
      auto issue_sdp = [&](int it) {
        const int s = it % kAbStages;
        const uint32_t base = smem_u32(smem_in + s * kAbStageBytes);
        { AB_T(t0); mbar_wait(&full_bar[s], (it / kAbStages) & 1); AB_ADD(0, t0); }
        { AB_T(t0); mbar_wait(sdp_free, (it & 1) ^ 1); AB_ADD(1, t0); }   // the previous item's S / dP have been read
        tc_fence_after();
        const uint64_t aq = umma_desc_k_sw128(base), bk = umma_desc_k_sw128(base + kInTile);
        const uint64_t ao = umma_desc_k_sw128(base + 3 * kInTile), bv = umma_desc_k_sw128(base + 2 * kInTile);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem_base + kColS, aq + 2 * k, bk + 2 * k, idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem_base + kColdP, ao + 2 * k, bv + 2 * k, idesc_s, k != 0);
        umma_commit(sdp_full);
      };
      // second phase of item `it`: dV = Pᵀ·dO, dK = dSᵀ·Q, dQ = dS·K
      auto issue_out = [&](int it) {
        const int s = it % kAbStages;
        const uint32_t base = smem_u32(smem_in + s * kAbStageBytes);
        const uint32_t qa = base, ka = base + kInTile, oa = base + 3 * kInTile;
        tc_fence_after();
        // blocks of 16 rows that hold no token of any sample are skipped (their P / dS rows and columns are zero)
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {            // 16 queries per instruction
          if ((G == 2 ? (kk & 3) : kk) >= NB) continue;
          const uint64_t ap = ab_desc_mn(pa + kk * 2048, kAbTile);
          const uint64_t ad = ab_desc_mn(da + kk * 2048, kAbTile);
          umma_f16(tmem_base + kColdV, ap, ab_desc_mn(oa + kk * 2048, 16), idesc_t, kk != 0);
          umma_f16(tmem_base + kColdK, ad, ab_desc_mn(qa + kk * 2048, 16), idesc_t, kk != 0);
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {            // 16 keys per instruction
          if ((G == 2 ? (kk & 3) : kk) >= NB) continue;
          const uint64_t ad = umma_desc_k_sw128(da + (kk >> 2) * kAbTile) + 2 * (kk & 3);
          umma_f16(tmem_base + kColdQ, ad, ab_desc_mn(ka + kk * 2048, 16), idesc_q, kk != 0);
        }
        umma_commit(out_full);
        umma_commit(&empty_bar[s]);                 // Q, K, V, dO of this item are no longer needed
      };
      // Issue order decided at run time: the second phase of item i as soon as its P / dS are in shared memory, the
      // first phase of item i+1 as soon as its operands have landed — whichever is ready first.  (In program order
      // "first phase of i+1, then second phase of i" the second phase waited for a load that, with two input stages,
      // can only start when the second phase of i−1 has completed; the other way round the soft-max warpgroup idled
      // through every second phase.)
      int next_sdp = 0, next_out = 0;
      uint32_t idle_polls = 0;
      long long t_idle = 0;
      while (next_out < n_mine) {
        if (next_out < next_sdp && mbar_try_wait(pds_ready, next_out & 1) && mbar_try_wait(out_free, (next_out & 1) ^ 1)) {
          AB_INC(12);
          issue_out(next_out);
          ++next_out;
          idle_polls = 0;
        } else if (next_sdp < n_mine && next_sdp <= next_out + 1 &&
                   mbar_try_wait(&full_bar[next_sdp % kAbStages], (next_sdp / kAbStages) & 1) &&
                   mbar_try_wait(sdp_free, (next_sdp & 1) ^ 1)) {
          issue_sdp(next_sdp);
          ++next_sdp;
          idle_polls = 0;
        } else if ((++idle_polls & 1023u) == 0) {   // bounded like mbar_wait: a pipeline bug must trap, not hang the GPU
          const long long now = clock64();
          if (idle_polls == 1024u) t_idle = now;
          else if (now - t_idle > 8000000000LL) {
            printf("gripb200: attention backward MMA thread stalled (block %d)\n", blockIdx.x);
            __trap();
          }
        }
      }
      AB_ADD(4, t_mma);
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================== soft-max / dS warpgroup (thread = query row) =====================
    reg_alloc<224>();
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int j = G == 2 ? r >> 6 : 0;
    const int qi = G == 2 ? r & 63 : r;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t t_s = tmem_base + lane_off + kColS + (G == 2 ? j * 64 : 0);
    const uint32_t t_dp = tmem_base + lane_off + kColdP + (G == 2 ? j * 64 : 0);
    const uint32_t p_row = smem_u32(smem_p) + r * 128;
    const uint32_t ds_row = smem_u32(smem_ds) + r * 128;
    const bool causal = p.causal != 0;
    const bool row_ok = qi < L;                     // padded query rows contribute nothing
    const int k_end = row_ok ? (causal ? min(L, qi + 1) : L) : 0;
    int it = 0;
    AB_T(t_wg1);
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
      { AB_T(t0); mbar_wait(sdp_full, it & 1); if (threadIdx.x == 128) AB_ADD(5, t0); }
      tc_fence_after();
      uint32_t v[kCols];   // raw scores → log2-domain scores → exponentials → probabilities, in place
      uint32_t dp[kCols];
      AB_T(t_ld);
#pragma unroll
      for (int c = 0; c < NB; ++c) ab_tmem_ld16(t_s + 16 * c, v + 16 * c);
#pragma unroll
      for (int c = 0; c < NB; ++c) ab_tmem_ld16(t_dp + 16 * c, dp + 16 * c);
      tmem_ld_wait();
      if (threadIdx.x == 128) AB_ADD(13, t_ld);
      AB_T(t_math);
      tc_fence_before();
      mbar_arrive(sdp_free);       // S and dP are in registers: the next item's may be issued
      float delta = 0.f;           // δ = Σ_j P_ij·dP_ij
      {
        // scores in log2 units: (q·k / 8)·log2(e).  The row reductions run on four interleaved partial accumulators:
        // a warpgroup has ONE warp per scheduler, so a single 64- or 80-long dependent chain of max / add / fma is
        // pure latency (the soft-max math was 3.6 k of the 5.4 k clocks an item took).
        constexpr float kScale = 0.18033688011112042f;   // > 0: the arg-max is taken on the raw scores
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int kj = 0; kj < kCols; ++kj) {
          const float x = kj < k_end ? __uint_as_float(v[kj]) : -INFINITY;
          v[kj] = __float_as_uint(x);
          m4[kj & 3] = fmaxf(m4[kj & 3], x);
        }
        float mxs = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * kScale;
        if (!row_ok) mxs = 0.f;
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kj = 0; kj < kCols; ++kj) {
          const float ex = fast_exp2(fmaf(__uint_as_float(v[kj]), kScale, -mxs));
          v[kj] = __float_as_uint(ex);
          s4[kj & 3] += ex;
        }
        const float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
        const float inv = row_ok ? 1.0f / sum : 0.f;
        float d4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kj = 0; kj < kCols; ++kj) {
          const float pj = __uint_as_float(v[kj]) * inv;          // P
          v[kj] = __float_as_uint(pj);
          d4[kj & 3] = fmaf(pj, __uint_as_float(dp[kj]), d4[kj & 3]);
        }
        delta = (d4[0] + d4[1]) + (d4[2] + d4[3]);
      }
      if (threadIdx.x == 128) AB_ADD(14, t_math);
      // the previous item's second phase must have finished reading the P / dS tiles (what a row does not write
      // below was zeroed at start-up and stays zero: nothing else ever writes these tiles)
      { AB_T(t0); mbar_wait(out_full, (it & 1) ^ 1); if (threadIdx.x == 128) AB_ADD(6, t0); }
      AB_T(t_st);
#pragma unroll
      for (int c8l = 0; c8l < 2 * NB; ++c8l) {                    // 8-key chunks of this row
        uint4 op, od;
        __half2* hp = reinterpret_cast<__half2*>(&op);
        __half2* hd = reinterpret_cast<__half2*>(&od);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int kj = 8 * c8l + 2 * t;
          const float p0 = __uint_as_float(v[kj]), p1 = __uint_as_float(v[kj + 1]);
          hp[t] = __floats2half2_rn(p0, p1);
          hd[t] = __floats2half2_rn(p0 * (__uint_as_float(dp[kj]) - delta), p1 * (__uint_as_float(dp[kj + 1]) - delta));
        }
        const uint32_t atom = (G == 2 ? j : c8l >> 3) * kAbTile;
        const uint32_t off = atom + (((c8l & 7) ^ (r & 7)) << 4);
        sts128(p_row + off, op);
        sts128(ds_row + off, od);
      }
      fence_proxy_async();         // generic-proxy writes of P, dS → visible to the tensor core
      mbar_arrive(pds_ready);
      if (threadIdx.x == 128) AB_ADD(15, t_st);
    }
    if (threadIdx.x == 128) AB_ADD(7, t_wg1);
  } else if (warp >= 8) {
    // ===================== output warpgroup (thread = output row) =====================
    reg_alloc<224>();
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const bool leader = threadIdx.x == 256;
    int it = 0;
    uint32_t n_store = 0;   // output tiles staged so far: tile n uses staging buffer n & 1
    AB_T(t_wg2);
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
      const int grp = item / H, h = item % H;
      const int n_s = min(G, p.B - grp * G);
      { AB_T(t0); mbar_wait(out_full, it & 1); if (leader) AB_ADD(8, t0); }
      tc_fence_after();
      uint32_t ov[3][64];
#pragma unroll
      for (int m = 0; m < 3; ++m)
        tmem_ld_32x64(tmem_base + lane_off + (m == 0 ? kColdQ : (m == 1 ? kColdK : kColdV)), ov[m]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(out_free);       // dQ, dK, dV are in registers: the next item's second phase may be issued
#pragma unroll
      for (int m = 0; m < 3; ++m, ++n_store) {
        uint8_t* stg = smem_stg + (n_store & 1) * kAbTile;
        // the store that last used this staging buffer (two tiles ago) must have read it
        if (leader) { AB_T(t0); tma_store_wait_read<1>(); AB_ADD(9, t0); }
        ab_wg_barrier(2);
        const float sc = m == 2 ? 1.0f : 0.125f;
        const uint32_t o_row = smem_u32(stg) + r * 128;
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
          uint4 o;
          __half2* hh = reinterpret_cast<__half2*>(&o);
#pragma unroll
          for (int t = 0; t < 4; ++t)
            hh[t] = __floats2half2_rn(__uint_as_float(ov[m][8 * c8 + 2 * t]) * sc, __uint_as_float(ov[m][8 * c8 + 2 * t + 1]) * sc);
          sts128(o_row + ((c8 ^ (r & 7)) << 4), o);
        }
        fence_proxy_async();
        ab_wg_barrier(2);
        if (leader) {
          for (int jj = 0; jj < n_s; ++jj)
            tma_store_2d(&tmDQKV, stg + jj * 8192, m * D + h * 64, (grp * G + jj) * L);
          tma_store_commit();
        }
      }
    }
    if (leader) AB_ADD(10, t_wg2);
    if (leader) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int G, int NB>
int launch_bwd_tc(gb_ctx* c, const CUtensorMap& tmQKV, const CUtensorMap& tmDO, const CUtensorMap& tmDQKV,
                  const AttnBwdParams& p, cudaStream_t st) {
  static bool done[16] = {false};
  if (!done[c->device & 15]) {
    GB_CUDA(c, cudaFuncSetAttribute(attn_bwd_tc_kernel<G, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, AbGeom<G, NB>::kSmem));
    done[c->device & 15] = true;
  }
  const int sms = gb_gemm_sms(c);
  const int grid = p.n_items < sms ? p.n_items : sms;
  attn_bwd_tc_kernel<G, NB><<<grid, kAbThreads, AbGeom<G, NB>::kSmem, st>>>(tmQKV, tmDO, tmDQKV, p);
  GB_LAUNCH_CHECK(c);
  return GB_OK;
}

}  // namespace

int gb_launch_attn_bwd_tc(gb_ctx* c, const void* qkv, const void* dout, void* dqkv, int B, int L, int D, int causal,
                          cudaStream_t st) {
  if (B <= 0) return GB_OK;
  if (L < 1 || L > 96 || D % 64 != 0)
    return gb_fail(c, GB_ERR_ARG, "attention backward (tcgen05): L=%d (1..96) D=%d (multiple of 64) unsupported", L, D);
  if ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(dqkv)) & 15)
    return gb_fail(c, GB_ERR_ARG, "attention backward: pointers must be 16-byte aligned");
  const int G = L <= 64 ? 2 : 1;
  CUtensorMap tmQKV, tmDO, tmDQKV;
  int rc = gb_make_tmap_2d_f16(c, &tmQKV, qkv, (uint64_t)B * L, (uint64_t)3 * D, (uint64_t)3 * D, (uint32_t)L);
  if (rc) return rc;
  rc = gb_make_tmap_2d_f16(c, &tmDO, dout, (uint64_t)B * L, (uint64_t)D, (uint64_t)D, (uint32_t)L);
  if (rc) return rc;
  rc = gb_make_tmap_2d_f16(c, &tmDQKV, dqkv, (uint64_t)B * L, (uint64_t)3 * D, (uint64_t)3 * D, (uint32_t)L);
  if (rc) return rc;
  AttnBwdParams p;
  p.B = B; p.L = L; p.H = D / 64; p.D = D; p.causal = causal;
  p.n_groups = (B + G - 1) / G;
  p.n_items = p.n_groups * p.H;
  switch ((L + 15) / 16) {
    case 1: return launch_bwd_tc<2, 1>(c, tmQKV, tmDO, tmDQKV, p, st);
    case 2: return launch_bwd_tc<2, 2>(c, tmQKV, tmDO, tmDQKV, p, st);
    case 3: return launch_bwd_tc<2, 3>(c, tmQKV, tmDO, tmDQKV, p, st);
    case 4: return launch_bwd_tc<2, 4>(c, tmQKV, tmDO, tmDQKV, p, st);
    case 5: return launch_bwd_tc<1, 5>(c, tmQKV, tmDO, tmDQKV, p, st);
    default: return launch_bwd_tc<1, 6>(c, tmQKV, tmDO, tmDQKV, p, st);
  }
}

#ifdef GB_ATTN_STALLS
extern "C" int gb_debug_attn_stalls(unsigned long long* out16, int reset) {
  if (cudaDeviceSynchronize() != cudaSuccess) return GB_ERR_CUDA;
  if (out16 && cudaMemcpyFromSymbol(out16, g_attn_stalls, 128) != cudaSuccess) return GB_ERR_CUDA;
  if (reset) {
    unsigned long long z[16] = {0};
    if (cudaMemcpyToSymbol(g_attn_stalls, z, 128) != cudaSuccess) return GB_ERR_CUDA;
  }
  return GB_OK;
}
#endif
