// Multi-head self-attention forward on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), for the short CLIP
// sequences (vision L = 50+P, text L ≤ 77; head dim 64):  O = softmax(Q·Kᵀ / 8 [+ causal mask]) · V  per (sample, head).
//
// Reference: nn.MultiheadAttention inside clip.model.ResidualAttentionBlock (need_weights=False; text tower:
// additive −inf upper-triangular mask), reached from models/clip_encoders.py:75-84,186.
//
// One work item = one head of G samples (G = 2 while L ≤ 64, else 1): sample j of the item occupies rows
// [64·j, 64·j + L) of three 128-row shared-memory tiles Q, K, V (128-byte rows, 128B swizzle — written by TMA boxes
// of exactly L rows straight out of the packed in-proj output qkv[B·L, 3·D]; the rows in between stay zero).
//   S = Q·Kᵀ         one 128×128×64 tcgen05.mma chain into TMEM; the blocks off the diagonal (query of one
//                     sample × keys of the other) are never read
//   P = softmax(S)    thread = query row: one tcgen05.ld of the row's own 64 (G = 2) or 128 (G = 1) key columns,
//                     soft-max in registers in log2 units (no cross-thread reduction), fp16 P written into a
//                     K-major swizzled 128×128 tile whose off-diagonal quadrants are zero
//   O = P·V           one 128×64×128 chain; V is consumed in place as an MN-major B operand (keys = K dimension)
//   O → fp16 → swizzled staging tile → TMA store of L rows per sample into a[B·L, D]
// Persistent CTAs, warp specialised: warp 0 TMA producer (3-stage ring of Q|K|V: the loads run two items ahead of
// the P·V that frees a stage — with two stages the next item's load could only start when the previous item was
// completely done and the kernel ran at the latency of one load per item), warp 1 MMA issuer, warp 2 TMEM
// allocator, two soft-max warpgroups that take alternate items (each owns its S and O accumulators and its P
// tile), so one group's soft-max overlaps the other's MMAs and stores.  The fp16 output tile is staged in the first
// atom of the group's own P tile (P is dead once P·V has completed), which is what leaves room for the third stage.
// Samples are visited from the last one down (see attn.cu: the in-proj GEMM's most recent rows are still in L2).
// Algorithmic HBM traffic: 8·D bytes per token (read q, k, v, write o); no other global access.
#include <stdlib.h>

#include "common.cuh"
#include "ctx.h"

using namespace gb;

namespace {

constexpr int kAtStages = 3;
constexpr int kAtTile = 128 * 128;            // one 128-row × 64-half tile, bytes
constexpr int kAtStageBytes = 3 * kAtTile;    // Q | K | V
constexpr int kAtPBytes = 2 * kAtTile;        // P: two 64-key atoms of 128 rows
constexpr int kAtThreads = 128 + 2 * 128;     // TMA, MMA, TMEM-alloc, idle + two soft-max warpgroups
constexpr int kAtSmem = kAtStages * kAtStageBytes + 2 * kAtPBytes + 1024 + 256;

struct AttnTcParams {
  int B, L, H, D;
  int causal;
  int n_groups;   // ceil(B / G)
  int n_items;    // n_groups · H
};

// B operand read MN-major (rows of the tile = K dimension, 64 contiguous N elements per 128-byte row, 128B swizzle):
// groups of 8 K-rows are 1024 B apart (SBO); the N extent is one swizzle atom, so LBO is not consulted.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__device__ __forceinline__ void wg_barrier(int id) {
  asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory");
}

// 16 consecutive accumulator columns into 16 registers (no wait: issued back to back, one wait for the row)
__device__ __forceinline__ void at_tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// G samples per item; NB = ceil(L / 16): blocks of 16 key columns a query row looks at (the soft-max and the P·V chain
// skip the rest: at L = 66 that is 80 of 128 columns)
template <int G, int NB>
__global__ void __launch_bounds__(kAtThreads, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmO,
                   const AttnTcParams p) {
  constexpr int kCols = 16 * NB;              // key columns a query row looks at
  static_assert(G == 2 ? NB <= 4 : (NB > 4 && NB <= 8), "G = 2 packs samples of up to 64 rows");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_qkv = smem;                                        // [stage][Q|K|V][128][128 B]
  uint8_t* smem_p = smem_qkv + kAtStages * kAtStageBytes;          // [wg][atom][128][128 B]; atom 0 doubles as the O staging tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_p + 2 * kAtPBytes);
  uint64_t* full_bar = bars;                  // [kAtStages]
  uint64_t* empty_bar = bars + kAtStages;     // [kAtStages]
  uint64_t* sfull_bar = bars + 2 * kAtStages; // [2] S of warpgroup w complete
  uint64_t* pready_bar = sfull_bar + 2;       // [2] P of warpgroup w in shared memory (and S read)
  uint64_t* ofull_bar = pready_bar + 2;       // [2] O of warpgroup w complete
  uint64_t* odone_bar = ofull_bar + 2;        // [2] O of warpgroup w read out of TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(odone_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmO);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kAtStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int w = 0; w < 2; ++w) {
      mbar_init(&sfull_bar[w], 1);
      mbar_init(&pready_bar[w], 128);
      mbar_init(&ofull_bar[w], 1);
      mbar_init(&odone_bar[w], 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);   // per warpgroup: S 128 columns + O 64 columns, at w·256
    tmem_relinquish();
  }
  // rows a TMA box never writes (≥ L in each 64-row half) and P's off-diagonal quadrants must read as zero
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const int n16 = (kAtStages * kAtStageBytes + 2 * kAtPBytes) / 16;
    for (int i = threadIdx.x; i < n16; i += kAtThreads) z[i] = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async();   // … also for the async proxy (TMA writes, UMMA reads) that touches the tiles next
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();  // the out-proj GEMM behind us may set itself up while we drain

  const int L = p.L, H = p.H, D = p.D;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int it = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
        const int s = it % kAtStages;
        const uint32_t ph = (it / kAtStages) & 1;
        const int grp = p.n_groups - 1 - item / H;   // last samples first
        const int h = item % H;
        const int n_s = min(G, p.B - grp * G);       // samples present in this item
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_expect_tx(&full_bar[s], 3u * n_s * L * 128u);
        uint8_t* st = smem_qkv + s * kAtStageBytes;
        for (int j = 0; j < n_s; ++j) {
          const int row0 = (grp * G + j) * L;
          for (int m = 0; m < 3; ++m)
            tma_load_2d(st + m * kAtTile + j * 8192, &tmQKV, &full_bar[s], m * D + h * 64, row0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_f16(128, 128);
      constexpr uint32_t idesc_pv = umma_idesc_f16(128, 64) | (1u << 16);   // B operand MN-major
      auto issue_pv = [&](int j) {
        const int w = j & 1, n = j >> 1, s = j % kAtStages;
        mbar_wait(&pready_bar[w], n & 1);            // P is in shared memory, S has been read
        mbar_wait(&odone_bar[w], (n & 1) ^ 1);       // the previous O of this warpgroup has left TMEM
        tc_fence_after();
        const uint32_t pa = smem_u32(smem_p + w * kAtPBytes);
        const uint32_t va = smem_u32(smem_qkv + s * kAtStageBytes + 2 * kAtTile);
        const uint32_t tmem_o = tmem_base + w * 256 + 128;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {             // 16 keys per instruction; blocks without a key are skipped
          if ((G == 2 ? (kk & 3) : kk) >= NB) continue;
          const uint64_t adesc = umma_desc_k_sw128(pa + (kk >> 2) * kAtTile) + 2 * (kk & 3);
          const uint64_t bdesc = umma_desc_mn_sw128(va + kk * 2048);
          umma_f16(tmem_o, adesc, bdesc, idesc_pv, kk != 0);
        }
        umma_commit(&ofull_bar[w]);
        umma_commit(&empty_bar[s]);                  // Q, K, V of this item are no longer needed
      };
      int it = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
        const int s = it % kAtStages, w = it & 1;
        mbar_wait(&full_bar[s], (it / kAtStages) & 1);
        tc_fence_after();
        const uint32_t qa = smem_u32(smem_qkv + s * kAtStageBytes);
        const uint64_t adesc = umma_desc_k_sw128(qa);
        const uint64_t bdesc = umma_desc_k_sw128(qa + kAtTile);
        const uint32_t tmem_s = tmem_base + w * 256;
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem_s, adesc + 2 * k, bdesc + 2 * k, idesc_qk, k != 0);
        umma_commit(&sfull_bar[w]);
        if (it >= 1) issue_pv(it - 1);
      }
      if (it >= 1) issue_pv(it - 1);
    }
  } else if (warp >= 4) {
    // ===================== soft-max warpgroups =====================
    const int w = (warp - 4) >> 2;          // warpgroup: items it ≡ w (mod 2) of this CTA
    const int q = warp & 3;                 // TMEM lane quadrant
    const int r = q * 32 + lane;            // query row of the tile = TMEM lane
    const int j = G == 2 ? r >> 6 : 0;      // sample of the item this row belongs to
    const int qi = G == 2 ? r & 63 : r;     // position inside the sample
    const uint32_t t_s = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + w * 256 + (G == 2 ? j * 64 : 0);
    const uint32_t t_o = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + w * 256 + 128;
    const uint32_t p_row = smem_u32(smem_p + w * kAtPBytes) + r * 128;     // + atom·16 KB + swizzled chunk
    uint8_t* o_tile = smem_p + w * kAtPBytes;
    const uint32_t o_row = smem_u32(o_tile) + r * 128;
    const bool causal = p.causal != 0;
    int n = 0;
    int it = w;
    for (int item = blockIdx.x + w * gridDim.x; item < p.n_items; item += 2 * gridDim.x, it += 2, ++n) {
      const int grp = p.n_groups - 1 - item / H;
      const int h = item % H;
      const int n_s = min(G, p.B - grp * G);
      // ---- S → P ----
      mbar_wait(&sfull_bar[w], n & 1);
      tc_fence_after();
      uint32_t v[kCols];
#pragma unroll
      for (int c = 0; c < NB; ++c) at_tmem_ld16(t_s + 16 * c, v + 16 * c);
      tmem_ld_wait();
      // scores in log2 units: (q·k / 8)·log2(e); key kj is visible when kj < L (and kj ≤ qi under the causal mask)
      const int k_end = causal ? min(L, qi + 1) : L;
      // four interleaved partial accumulators: one warp per scheduler cannot hide a 64-long dependent max / add chain.
      // The arg-max is taken on the raw scores and the scale (1/8·log2 e > 0) goes into the exponent's FFMA.
      constexpr float kScale = 0.18033688011112042f;
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int kj = 0; kj < kCols; ++kj) {
        const float x = kj < k_end ? __uint_as_float(v[kj]) : -INFINITY;
        v[kj] = __float_as_uint(x);
        m4[kj & 3] = fmaxf(m4[kj & 3], x);
      }
      const float mxs = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * kScale;   // key 0 is always visible → finite
      float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int kj = 0; kj < kCols; ++kj) {
        const float e = fast_exp2(fmaf(__uint_as_float(v[kj]), kScale, -mxs));
        v[kj] = __float_as_uint(e);
        s4[kj & 3] += e;
      }
      const float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
      const float inv = 1.0f / sum;
      // P goes where this warpgroup's previous output tile was staged: its TMA stores must have read it
      if (threadIdx.x == 128 + w * 128) tma_store_wait_read<0>();
      wg_barrier(1 + w);
      // the staging tile covered atom 0 of every row: what this row does not write there below must be zero again
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {
        const bool mine = G == 2 ? (j == 0 && c8 < 2 * NB) : true;
        if (!mine) sts128(p_row + ((c8 ^ (r & 7)) << 4), make_uint4(0, 0, 0, 0));
      }
#pragma unroll
      for (int c8 = 0; c8 < kCols / 8; ++c8) {
        uint4 o;
        __half2* hh = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int t = 0; t < 4; ++t)
          hh[t] = __floats2half2_rn(__uint_as_float(v[8 * c8 + 2 * t]) * inv, __uint_as_float(v[8 * c8 + 2 * t + 1]) * inv);
        const int atom = G == 2 ? j : c8 >> 3;       // 64-key atom this chunk lives in
        sts128(p_row + atom * kAtTile + (((c8 & 7) ^ (r & 7)) << 4), o);
      }
      fence_proxy_async();       // generic-proxy writes of P → visible to the tensor core
      tc_fence_before();
      mbar_arrive(&pready_bar[w]);
      // ---- O → global ----
      mbar_wait(&ofull_bar[w], n & 1);
      tc_fence_after();
      uint32_t ov[64];
      tmem_ld_32x64(t_o, ov);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&odone_bar[w]);
      // P·V has completed (ofull), so the P tile is dead: its first atom is the staging tile
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {
        uint4 o;
        __half2* hh = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int t = 0; t < 4; ++t)
          hh[t] = __floats2half2_rn(__uint_as_float(ov[8 * c8 + 2 * t]), __uint_as_float(ov[8 * c8 + 2 * t + 1]));
        sts128(o_row + ((c8 ^ (r & 7)) << 4), o);
      }
      fence_proxy_async();
      wg_barrier(1 + w);
      if (threadIdx.x == 128 + w * 128) {
        for (int jj = 0; jj < n_s; ++jj)
          tma_store_2d(&tmO, o_tile + jj * 8192, h * 64, (grp * G + jj) * L);
        tma_store_commit();
      }
    }
    if (threadIdx.x == 128 + w * 128) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int G, int NB>
int launch_tc(gb_ctx* c, const CUtensorMap& tmQKV, const CUtensorMap& tmO, const AttnTcParams& p, cudaStream_t st) {
  static bool done[16] = {false};
  if (!done[c->device & 15]) {
    GB_CUDA(c, cudaFuncSetAttribute(attn_fwd_tc_kernel<G, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmem));
    done[c->device & 15] = true;
  }
  const int sms = gb_gemm_sms(c);
  const int grid = p.n_items < sms ? p.n_items : sms;
  attn_fwd_tc_kernel<G, NB><<<grid, kAtThreads, kAtSmem, st>>>(tmQKV, tmO, p);
  GB_LAUNCH_CHECK(c);
  return GB_OK;
}

}  // namespace

// GB_ATTN_LEGACY=1 keeps the mma.sync kernel of attn.cu (A/B measurements)
bool gb_attn_tc_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("GB_ATTN_LEGACY");
    on = (e && e[0] == '1') ? 0 : 1;
  }
  return on != 0;
}

int gb_launch_attn_fwd_tc(gb_ctx* c, const void* qkv, void* out, int B, int L, int D, int causal, cudaStream_t st) {
  if (B <= 0) return GB_OK;
  if (L < 1 || L > 128 || D % 64 != 0)
    return gb_fail(c, GB_ERR_ARG, "attention (tcgen05): L=%d (1..128) D=%d (multiple of 64) unsupported", L, D);
  if ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out)) & 15)
    return gb_fail(c, GB_ERR_ARG, "attention: pointers must be 16-byte aligned");
  const int G = L <= 64 ? 2 : 1;
  CUtensorMap tmQKV, tmO;
  int rc = gb_make_tmap_2d_f16(c, &tmQKV, qkv, (uint64_t)B * L, (uint64_t)3 * D, (uint64_t)3 * D, (uint32_t)L);
  if (rc) return rc;
  rc = gb_make_tmap_2d_f16(c, &tmO, out, (uint64_t)B * L, (uint64_t)D, (uint64_t)D, (uint32_t)L);
  if (rc) return rc;
  AttnTcParams p;
  p.B = B; p.L = L; p.H = D / 64; p.D = D; p.causal = causal;
  p.n_groups = (B + G - 1) / G;
  p.n_items = p.n_groups * p.H;
  switch ((L + 15) / 16) {
    case 1: return launch_tc<2, 1>(c, tmQKV, tmO, p, st);
    case 2: return launch_tc<2, 2>(c, tmQKV, tmO, p, st);
    case 3: return launch_tc<2, 3>(c, tmQKV, tmO, p, st);
    case 4: return launch_tc<2, 4>(c, tmQKV, tmO, p, st);
    case 5: return launch_tc<1, 5>(c, tmQKV, tmO, p, st);
    case 6: return launch_tc<1, 6>(c, tmQKV, tmO, p, st);
    case 7: return launch_tc<1, 7>(c, tmQKV, tmO, p, st);
    default: return launch_tc<1, 8>(c, tmQKV, tmO, p, st);
  }
}
