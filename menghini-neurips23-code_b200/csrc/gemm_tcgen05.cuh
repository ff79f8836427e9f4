// Persistent warp-specialised tcgen05 GEMM for sm_100a.
//
//   out[M,N] = epilogue( A[M,K] · W[N,K]ᵀ )        A, W fp16 K-major; fp32 accumulate in TMEM
//
// Roles (one CTA per SM, 384 threads):
//   warp 0   TMA producer   – cp.async.bulk.tensor (128B swizzle) into a kStages ring
//   warp 1   MMA issuer     – one elected lane issues tcgen05.mma (128×BN×16), commits to mbarriers
//   warp 2   TMEM allocator – 2×BN fp32 columns (double-buffered accumulator)
//   warps 4-11 epilogue     – tcgen05.ld (one accumulator row per thread, two warps per TMEM lane
//                             quadrant) → bias / QuickGELU / residual / fp32|fp16 store, overlapped
//                             with the next tile's MMAs
//
// This is the contraction behind every nn.Linear / conv1 / projection on the CLIP towers
// (reference call sites: third-party clip.model.ResidualAttentionBlock via
//  models/clip_encoders.py:75-84,186 ; SURVEY.md §2.1 O1,O4,O6,O7,O8,O9).
#pragma once
#include "common.cuh"

namespace gb {

// Optional stall accounting of the CTA-pair kernel (build with -DGB_GEMM_STALLS; read back with
// gb_debug_gemm_stalls): cycles the MMA thread waits for operands / for a drained accumulator, the
// producer for a free slot, and epilogue warp 4 for a finished accumulator.
#ifdef GB_GEMM_STALLS
__device__ unsigned long long g_gemm_stalls[16];
#define GB_STALL_DECL(v) long long v = 0
#define GB_STALL_T(t) const long long t = clock64()
#define GB_STALL_ADD(v, t) v += clock64() - t
#define GB_STALL_PUT(i, v) atomicAdd(&g_gemm_stalls[i], (unsigned long long)(v))
#else
#define GB_STALL_DECL(v)
#define GB_STALL_T(t)
#define GB_STALL_ADD(v, t)
#define GB_STALL_PUT(i, v)
#endif

constexpr int kBM = 128;   // rows per tile  (UMMA M)
constexpr int kBK = 64;    // fp16 elements per k-block = one 128 B swizzle atom
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 384;  // TMA, MMA, TMEM-alloc, idle + 8 epilogue warps

struct GemmParams {
  int M, N, K;
  void* out;            // fp16 (default) or fp32 (out_f32)
  int ldo;              // elements
  const float* bias;    // [N] or nullptr
  const __half* resid;  // [M,ldr] or nullptr; may alias out
  int ldr;
  int act;      // 0 none, 1 QuickGELU, 2 multiply by QuickGELU'(aux) (backward of act 1)
  int out_f32;  // 1 → fp32 store
  __half* aux;  // [M,ldo] fp16 or nullptr.  act 1: receives the pre-activation (tape for backward);
                // act 2: the saved pre-activation that is read.
  // LayerNorm folded into the contraction (CTA-pair kernel, fp16 output only).  With W' = W∘γ,
  // s_n = Σ_k W'[n,k] and b' = b + W·β:   LN(x)·Wᵀ + b = rstd_r·(x·W'ᵀ − μ_r·s_n) + b'_n,
  // so A is the raw residual stream and the row statistics are applied in the epilogue.
  const float* ln_stats;  // [M][2] = (μ·rstd, rstd) of each row of A, or nullptr
  const float* col_sum;   // s_n, [N]
  // … or the row statistics still in the form the producing GEMM left them (its stats_out): the
  // epilogue merges the ln_nparts segments of a row itself, one tile ahead (no finalize launch)
  const float4* ln_parts;  // [ln_nparts][M] (x0, Σ(x−x0), Σ(x−x0)², −), or nullptr
  int ln_nparts;           // ≤ kLnMaxParts
  float ln_seg_n;          // columns per segment
  // shifted partial sums of every OUTPUT row over this warp's 128 columns, [N/128][M] float4 =
  // (x0, Σ(x−x0), Σ(x−x0)², −) with x0 the segment's first stored value, or nullptr: the statistics the
  // next LayerNorm needs, produced where the rows are written.  The shift keeps the later variance
  // computation free of the E[x²]−μ² cancellation when a row's mean dwarfs its spread.
  float* stats_out;
};

template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kABytes = kBM * kBK * 2;  // 16 KB
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BN;  // power of two: 256 or 512
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

// Epilogue of one 128 x BN accumulator tile by the 8 epilogue warps of a CTA.  TMEM lane quadrant
// q = warp % 4 (hardware rule); the two warps of a quadrant split the BN columns.  The operand that
// has to come from global memory (residual, or the saved pre-activation for act 2) is prefetched one
// 32-column chunk ahead — the first chunk before the accumulator is even ready — so its latency hides
// behind the TMEM read-out of the previous one.  `wait_acc` blocks until the accumulator is complete.
template <int BN, typename WaitAcc>
__device__ __forceinline__ void gemm_epilogue_tile(const GemmParams& p, uint32_t tmem_acc, int m0,
                                                   int n_base, int warp, int lane,
                                                   WaitAcc&& wait_acc) {
  const int q = warp & 3;
  const int half = (warp - 4) >> 2;
  constexpr int kChunks = BN / 64;  // 32-column chunks per warp
  const __half* pre_base = (p.act == 2) ? p.aux : p.resid;
  const int pre_ld = (p.act == 2) ? p.ldo : p.ldr;
  const int n0 = n_base + half * (BN / 2);
  const int row = m0 + q * 32 + lane;
  const bool row_ok = row < p.M;
  const bool has_pre = pre_base != nullptr && row_ok;
  uint4 pre[4], pre_next[4];
  if (has_pre) {
    const uint4* r4 = reinterpret_cast<const uint4*>(pre_base + (size_t)row * pre_ld + n0);
#pragma unroll
    for (int j = 0; j < 4; ++j) pre[j] = r4[j];
  }
  wait_acc();
  tc_fence_after();
  const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(q * 32) << 16) + half * (BN / 2);
#pragma unroll 1
  for (int c = 0; c < kChunks; ++c) {
    uint32_t v[32];
    tmem_ld_32x32(taddr + c * 32, v);
    const int col0 = n0 + c * 32;
    if (has_pre && c + 1 < kChunks) {
      const uint4* r4 = reinterpret_cast<const uint4*>(pre_base + (size_t)row * pre_ld + col0 + 32);
#pragma unroll
      for (int j = 0; j < 4; ++j) pre_next[j] = r4[j];
    }
    tmem_ld_wait();
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
    if (p.bias != nullptr) {
      const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 bb = __ldg(b4 + j);
        f[4 * j + 0] += bb.x; f[4 * j + 1] += bb.y; f[4 * j + 2] += bb.z; f[4 * j + 3] += bb.w;
      }
    }
    if (p.act == 1) {
      if (p.aux != nullptr && row_ok) {
        uint4* a4 = reinterpret_cast<uint4*>(p.aux + (size_t)row * p.ldo + col0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 o;
          __half2* h = reinterpret_cast<__half2*>(&o);
#pragma unroll
          for (int t = 0; t < 4; ++t)
            h[t] = __floats2half2_rn(f[8 * j + 2 * t], f[8 * j + 2 * t + 1]);
          a4[j] = o;
        }
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = quick_gelu(f[j]);
    } else if (p.act == 2 && has_pre) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __half2* h = reinterpret_cast<const __half2*>(&pre[j]);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 x = __half22float2(h[t]);
          f[8 * j + 2 * t] *= quick_gelu_grad(x.x);
          f[8 * j + 2 * t + 1] *= quick_gelu_grad(x.y);
        }
      }
    }
    if (row_ok) {
      if (p.act != 2 && has_pre) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const __half2* h = reinterpret_cast<const __half2*>(&pre[j]);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 rf = __half22float2(h[t]);
            f[8 * j + 2 * t] += rf.x;
            f[8 * j + 2 * t + 1] += rf.y;
          }
        }
      }
      if (p.out_f32) {
        float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) +
                                               (size_t)row * p.ldo + col0);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          o4[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
      } else {
        uint4* o4 = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.out) +
                                             (size_t)row * p.ldo + col0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 o;
          __half2* h = reinterpret_cast<__half2*>(&o);
#pragma unroll
          for (int t = 0; t < 4; ++t)
            h[t] = __floats2half2_rn(f[8 * j + 2 * t], f[8 * j + 2 * t + 1]);
          o4[j] = o;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) pre[j] = pre_next[j];
  }
}

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_f16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA,
                        const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;

  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024 B alignment.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * Cfg::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;                    // [kStages]
  uint64_t* empty_bar = bars + kStages;         // [kStages]
  uint64_t* tfull_bar = bars + 2 * kStages;     // [2]
  uint64_t* tempty_bar = bars + 2 * kStages + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (p.M + kBM - 1) / kBM;
  const int n_tiles = p.N / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int k_blocks = p.K / kBK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 8);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above overlapped the predecessor's tail (programmatic dependent launch)
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * kBM;
        const int n0 = (tile % n_tiles) * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_2d(smem_a + stage * Cfg::kABytes, &tmA, &full_bar[stage], kb * kBK, m0);
          tma_load_2d(smem_b + stage * Cfg::kBBytes, &tmB, &full_bar[stage], kb * kBK, n0);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(kBM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t adesc = umma_desc_k_sw128(smem_u32(smem_a + stage * Cfg::kABytes));
          const uint64_t bdesc = umma_desc_k_sw128(smem_u32(smem_b + stage * Cfg::kBBytes));
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            // +32 B per UMMA_K step inside the 128 B swizzle atom (address field is in 16 B units)
            umma_f16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once the MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[as]);  // accumulator ready for the epilogue
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (8 warps) =====================
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int m0 = (tile / n_tiles) * kBM;
      const int n0 = (tile % n_tiles) * BN;
      gemm_epilogue_tile<BN>(p, tmem_base + as * BN, m0, n0, warp, lane,
                             [&]() { mbar_wait(&tfull_bar[as], aphase); });
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// Epilogue of one 128 x 256 accumulator tile with TMA stores (fp16 output).  Each of the 8 epilogue
// warps owns 32 rows x 128 columns.
//  * A tcgen05.ld issued while the pair's MMAs are running takes ≈1-2 k clocks to deliver, so the warp
//    pulls its entire 32 x 128 fp32 slice into registers with two back-to-back x64 loads, pays that
//    latency once per tile, and releases the TMEM accumulator to the MMA thread immediately — the
//    arithmetic, packing and stores then overlap the next-but-one tile's MMAs.  The epilogue
//    warpgroups run with 232 registers (setmaxnreg); the TMA / MMA warpgroup gives its registers up.
//  * The per-column constants (bias, and Σ_k W'[n,k] of a folded LayerNorm) of the warp's 128 columns
//    are fetched one tile ahead with one coalesced 16-byte load per lane and parked in a warp-private
//    shared-memory strip; the chunk loop reads them back as broadcasts.  (Read straight from global
//    memory they cost an L2 round trip per 32-column chunk — the top stall of the epilogue.)
//  * Output: 32 columns at a time into a private 32 x 64-byte shared-memory half-slab (64B-swizzled:
//    quarter-warp accesses are conflict-free); one lane hands it to the TMA engine, which writes whole
//    row segments and clips rows >= M.  Direct per-thread stores (32 row-strided 16-byte pieces per
//    instruction) were the bottleneck of every K=768 GEMM.  Two half-slabs per warp alternate, so one
//    is rewritten only after the store issued two chunks earlier has finished reading it.
#ifndef GB_SLAB_COLS
#define GB_SLAB_COLS 32  // columns per TMA store of the epilogue: 32 (2 KB half-slabs, 64B swizzle) or 64 (4 KB, 128B)
#endif
#ifndef GB_GELU_SHARE
#define GB_GELU_SHARE 4  // QuickGELUs per shared reciprocal (1, 2 or 4)
#endif
#ifndef GB_GELU_H2
#define GB_GELU_H2 1  // CTA-pair kernel: QuickGELU on the fp16-rounded pre-activation in packed fp16 (quick_gelu_h2)
#endif

// Epilogue geometry of the CTA-pair kernel: GB_EPI_WARPS epilogue warps (8 or 16) — TMEM lane quadrant
// warp % 4, column group (warp − 4) / 4 of 256 / (GB_EPI_WARPS / 4) columns.  Measured on B200 (same box,
// sustained loops): 16 warps of 112 registers are SLOWER than 8 warps of 232 on every flavour (plain fc
// 1029 vs 1084 TFLOP/s; LN+GELU fc in the ViT loop 395 vs 270 µs — the 12 statistics segments per row and the
// 64-column slices no longer fit the register budget), so 8 is the default.
#ifndef GB_EPI_WARPS
#define GB_EPI_WARPS 8
#endif
constexpr int kEpiWarps2 = GB_EPI_WARPS;
static_assert(kEpiWarps2 == 8 || kEpiWarps2 == 16, "8 or 16 epilogue warps");
constexpr int kEpiCols = 256 / (kEpiWarps2 / 4);
constexpr int kGemm2Threads = 128 + 32 * kEpiWarps2;
constexpr int kEpiRegs = kEpiWarps2 == 8 ? 232 : 112;   // setmaxnreg budgets: (128·kCtlRegs + 32·warps·kEpiRegs)
constexpr int kCtlRegs = kEpiWarps2 == 8 ? 40 : 24;     //   must fit 65536 / threads rounded down to 8, per thread
constexpr int kLnMaxParts = 768 / kEpiCols;             // statistics segments of a row (D ≤ 768)

// (μ·rstd, rstd) of one row from its per-segment shifted partials (x0, Σ(x−x0), Σ(x−x0)²): per segment
// mean and centred second moment, then Chan's pairwise update in a fixed segment order (deterministic,
// free of the E[x²]−μ² cancellation).
__device__ __forceinline__ float2 ln_merge_parts(const float4 (&lp)[kLnMaxParts], int nparts, float seg_n) {
  // All segments have the same length, so Chan's weights are constants of the segment index: after i
  // segments  mean += δ/(i+1)  and  M2 += M2_seg + δ²·seg_n·i/(i+1).  No run-time divisions (the IEEE
  // division sequence of the straightforward form was a sixth of the LN-folding epilogues' instructions);
  // 1/seg_n is exact (seg_n is a power of two).
  const float inv_seg = 1.0f / seg_n;
  float mean = 0.f, m2 = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxParts; ++i) {
    if (i < nparts) {
      const float4 t = lp[i];
      const float sm = t.y * inv_seg;              // segment mean − x0
      const float mp = t.x + sm;                   // segment mean
      const float m2p = fmaf(-sm, t.y, t.z);       // Σ (x − segment mean)²
      const float delta = mp - mean;
      mean = fmaf(delta, 1.0f / (float)(i + 1), mean);
      m2 += m2p + delta * delta * (seg_n * ((float)i / (float)(i + 1)));
    }
  }
  const float rstd = rsqrtf(fmaxf(m2 * (1.0f / ((float)nparts * seg_n)), 0.f) + 1e-5f);
  return make_float2(mean * rstd, rstd);
}
__device__ __forceinline__ void ln_load_parts(const GemmParams& p, int row, float4 (&lp)[kLnMaxParts]) {
#pragma unroll
  for (int i = 0; i < kLnMaxParts; ++i)
    if (i < p.ln_nparts) lp[i] = __ldg(p.ln_parts + (size_t)i * p.M + row);
}

// Epilogue flavours of the CTA-pair kernel: one kernel instantiation each, so that an instantiation's
// (fully unrolled) epilogue stays small — with every flavour behind run-time flags in one kernel the
// epilogue warps spent 15-25 % of their issue slots waiting for instruction fetches.
enum : int {
  kEpiPlain = 0,   // [+ bias]
  kEpiLn = 1,      // folded LayerNorm [+ bias]
  kEpiResid = 2,   // [+ bias] + residual [→ row statistics]
  kEpiGelu = 3,    // [+ bias], QuickGELU [pre-activation → aux]
  kEpiLnGelu = 4,  // folded LayerNorm [+ bias], QuickGELU [pre-activation → aux]
  kEpiAct2 = 5,    // × QuickGELU'(aux)   (backward of the activation)
  kEpiModes = 6
};

template <int kMode, typename WaitAcc, typename ReleaseAcc>
__device__ __forceinline__ void gemm_epilogue_tile_tma(const GemmParams& p, const CUtensorMap* tmC,
                                                       uint8_t* slabs, uint32_t tmem_acc, int m0,
                                                       int n_base, int warp, int lane, float2& ln_st,
                                                       int next_m0, uint32_t consts_s, uint32_t consts_next_s,
                                                       int next_n_base, uint4 (&pre_lo)[kEpiCols / 16],
                                                       float4 (&lp)[kLnMaxParts], bool& lp_pending,
                                                       WaitAcc&& wait_acc, ReleaseAcc&& release_acc) {
  constexpr int kCols = kEpiCols;         // accumulator columns per warp: 128 (8 warps) or 64 (16 warps)
  constexpr int kChunks = kCols / 32;     // 32-column chunks, one half-slab each
  constexpr int kPreHalf = kCols / 16;    // 16-byte pieces of fp16 in half of a warp's row slice
  constexpr bool kLn = kMode == kEpiLn || kMode == kEpiLnGelu;
  constexpr bool kGelu = kMode == kEpiGelu || kMode == kEpiLnGelu;
  constexpr bool kAct2 = kMode == kEpiAct2;
  constexpr bool kPre = kMode == kEpiResid || kMode == kEpiAct2;
  const int q = warp & 3;
  const int cg = (warp - 4) >> 2;   // column group of this warp
  const __half* pre_base = kAct2 ? p.aux : p.resid;
  const int pre_ld = kAct2 ? p.ldo : p.ldr;
  const int n0 = n_base + cg * kCols;
  const int row0 = m0 + q * 32;
  const int row = row0 + lane;
  const bool row_ok = row < p.M;
  const bool has_pre = kPre && pre_base != nullptr && row_ok;
  // The operand that comes from global memory (residual, or the saved pre-activation for act 2) is
  // fetched ahead: the first half of this row's columns was requested at the end of the previous tile
  // (pre_lo, carried across tiles); the second half is requested now, before the accumulator is even ready.
  uint4 pre_hi[kPre ? kPreHalf : 1];
  if (has_pre) {
    const uint4* r4 = reinterpret_cast<const uint4*>(pre_base + (size_t)row * pre_ld + n0 + kCols / 2);
#pragma unroll
    for (int j = 0; j < kPreHalf; ++j) pre_hi[j] = r4[j];
  }
  // row statistics of the folded LayerNorm: (μ·rstd, rstd) of this tile's row were fetched while the
  // previous tile was processed; the next tile's are requested now
  // (every load that is still in flight when a chunk's fence.proxy.async executes stalls the warp there —
  // the fence comes with a CTA-wide memory barrier — so look-ahead loads are issued after the tile's last
  // fence and land while the warp waits for the next accumulator: the statistics partials requested at
  // the end of the previous tile are merged here)
  if constexpr (kLn) {
    if (lp_pending) ln_st = ln_merge_parts(lp, p.ln_nparts, p.ln_seg_n);
  }
  const float ln_mr = -ln_st.x, ln_rstd = ln_st.y;
  float2 ln_next = make_float2(0.f, 1.f);
  const int ln_nrow = next_m0 + q * 32 + lane;
  const bool ln_parts_next = kLn && p.ln_parts != nullptr && next_m0 >= 0 && ln_nrow < p.M;
  if (kLn && p.ln_parts == nullptr && next_m0 >= 0 && ln_nrow < p.M)
    ln_next = *reinterpret_cast<const float2*>(p.ln_stats + (size_t)ln_nrow * 2);
  // next tile's per-column constants: lane l fetches kCols/32 consecutive columns of the warp's slice
  constexpr int kCPL = kCols / 32;
  float nb[kCPL], ns[kCPL];
#pragma unroll
  for (int i = 0; i < kCPL; ++i) nb[i] = ns[i] = 0.f;
  if (!kAct2 && next_n_base >= 0) {
    const int nc = next_n_base + cg * kCols + kCPL * lane;
#pragma unroll
    for (int i = 0; i < kCPL; ++i) {
      if (p.bias != nullptr) nb[i] = __ldg(p.bias + nc + i);
      if (kLn) ns[i] = __ldg(p.col_sum + nc + i);
    }
  }
  float st_sum = 0.f, st_sq = 0.f, st_x0 = 0.f;
  wait_acc();
  tc_fence_after();
  const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(q * 32) << 16) + cg * kCols;
  uint32_t v[kCols];
#pragma unroll
  for (int i = 0; i < kCols / 64; ++i) tmem_ld_32x64(taddr + 64 * i, v + 64 * i);
  tmem_ld_wait();
  tc_fence_before();
  __syncwarp();
  release_acc();  // the accumulator is in registers: the MMA thread may overwrite it
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {  // 32-column chunks, one half-slab each
    const int col0 = n0 + c * 32;
    uint4 pre[kPre ? 4 : 1];
    if constexpr (kPre) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        pre[j] = c < kChunks / 2 ? pre_lo[4 * c + j] : pre_hi[4 * (c - kChunks / 2) + j];
    }
#if GB_SLAB_COLS == 64
    uint8_t* slab = slabs + ((c >> 1) & 1) * 4096;
    const uint32_t slab_s = smem_u32(slab);
    if ((c & 1) == 0) {
      // the store that last used this slab (two stores ago) must be done reading it
      if (lane == 0) tma_store_wait_read<1>();
      __syncwarp();
    }
#else
    uint8_t* slab = slabs + (c & 1) * 2048;
    const uint32_t slab_s = smem_u32(slab);
    // the store that last used this half-slab (two chunks ago) must be done reading it
    if (lane == 0) tma_store_wait_read<1>();
    __syncwarp();
#endif
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[32 * c + j]);
    if constexpr (kLn) {
      // rstd·(acc − μ·s_n) + b_n = acc·rstd + (b_n − (μ·rstd)·s_n): two FMAs per element
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 bb = lds128f(consts_s + (c * 32 + 4 * j) * 4);
        const float4 ss = lds128f(consts_s + kCols * 4 + (c * 32 + 4 * j) * 4);
        f[4 * j + 0] = fmaf(f[4 * j + 0], ln_rstd, fmaf(ln_mr, ss.x, bb.x));
        f[4 * j + 1] = fmaf(f[4 * j + 1], ln_rstd, fmaf(ln_mr, ss.y, bb.y));
        f[4 * j + 2] = fmaf(f[4 * j + 2], ln_rstd, fmaf(ln_mr, ss.z, bb.z));
        f[4 * j + 3] = fmaf(f[4 * j + 3], ln_rstd, fmaf(ln_mr, ss.w, bb.w));
      }
    } else if (!kAct2 && p.bias != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 bb = lds128f(consts_s + (c * 32 + 4 * j) * 4);
        f[4 * j + 0] += bb.x; f[4 * j + 1] += bb.y; f[4 * j + 2] += bb.z; f[4 * j + 3] += bb.w;
      }
    }
    uint32_t act_h2[(kGelu && GB_GELU_H2) ? 16 : 1];   // packed fp16 activations (GB_GELU_H2)
    if constexpr (kGelu) {
#if GB_GELU_H2
      // pre-activation → fp16 once (it is what the tape keeps and what the reference's fp16 path feeds the
      // activation), QuickGELU in packed fp16
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 o;
        __half2* h = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int t = 0; t < 4; ++t)
          h[t] = __floats2half2_rn(f[8 * j + 2 * t], f[8 * j + 2 * t + 1]);
        if (p.aux != nullptr && row_ok)
          reinterpret_cast<uint4*>(p.aux + (size_t)row * p.ldo + col0)[j] = o;
        act_h2[4 * j + 0] = quick_gelu_h2(o.x);
        act_h2[4 * j + 1] = quick_gelu_h2(o.y);
        act_h2[4 * j + 2] = quick_gelu_h2(o.z);
        act_h2[4 * j + 3] = quick_gelu_h2(o.w);
      }
#else
      if (p.aux != nullptr && row_ok) {
        uint4* a4 = reinterpret_cast<uint4*>(p.aux + (size_t)row * p.ldo + col0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 o;
          __half2* h = reinterpret_cast<__half2*>(&o);
#pragma unroll
          for (int t = 0; t < 4; ++t)
            h[t] = __floats2half2_rn(f[8 * j + 2 * t], f[8 * j + 2 * t + 1]);
          a4[j] = o;
        }
      }
#if GB_GELU_SHARE == 4
#pragma unroll
      for (int j = 0; j < 8; ++j) quick_gelu4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
#elif GB_GELU_SHARE == 2
#pragma unroll
      for (int j = 0; j < 16; ++j) quick_gelu2(f[2 * j], f[2 * j + 1]);
#else
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = quick_gelu(f[j]);
#endif
#endif
    } else if (kAct2 && has_pre) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __half2* h = reinterpret_cast<const __half2*>(&pre[j]);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 x = __half22float2(h[t]);
          f[8 * j + 2 * t] *= quick_gelu_grad(x.x);
          f[8 * j + 2 * t + 1] *= quick_gelu_grad(x.y);
        }
      }
    }
    if (kMode == kEpiResid && has_pre) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __half2* h = reinterpret_cast<const __half2*>(&pre[j]);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 rf = __half22float2(h[t]);
          f[8 * j + 2 * t] += rf.x;
          f[8 * j + 2 * t + 1] += rf.y;
        }
      }
    }
    // 64 bytes of this thread's row → the four 16-byte pieces of half-slab row `lane`
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 o;
      __half2* h = reinterpret_cast<__half2*>(&o);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        if constexpr (kGelu && GB_GELU_H2) reinterpret_cast<uint32_t*>(&o)[t] = act_h2[4 * j + t];
        else h[t] = __floats2half2_rn(f[8 * j + 2 * t], f[8 * j + 2 * t + 1]);
        if (kMode == kEpiResid && p.stats_out != nullptr) {  // statistics of the values as stored (fp16-rounded)
          const float2 r = __half22float2(h[t]);
          if (c == 0 && j == 0 && t == 0) st_x0 = r.x;
          const float d0 = r.x - st_x0, d1 = r.y - st_x0;
          st_sum += d0 + d1;
          st_sq = fmaf(d0, d0, fmaf(d1, d1, st_sq));
        }
      }
#if GB_SLAB_COLS == 64
      sts128(slab_s + lane * 128 + ((((c & 1) * 4 + j) ^ (lane & 7)) << 4), o);
#else
      sts128(slab_s + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4), o);
#endif
    }
#if GB_SLAB_COLS == 64
    if (c & 1) {
      fence_proxy_async();  // generic-proxy writes → visible to the TMA engine
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(tmC, slab, col0 - 32, row0);
        tma_store_commit();
      }
    }
#else
    fence_proxy_async();  // generic-proxy writes → visible to the TMA engine
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(tmC, slab, col0, row0);
      tma_store_commit();
    }
#endif
  }
  if (!kAct2 && next_n_base >= 0) {  // park the next tile's constants (all reads of this tile's are done)
#pragma unroll
    for (int i = 0; i < kCPL; ++i) {
      sts32f(consts_next_s + (kCPL * lane + i) * 4, nb[i]);
      if (kLn) sts32f(consts_next_s + kCols * 4 + (kCPL * lane + i) * 4, ns[i]);
    }
    __syncwarp();
  }
  if (kMode == kEpiResid && p.stats_out != nullptr && row_ok)
    *reinterpret_cast<float4*>(p.stats_out + ((size_t)(n0 / kCols) * p.M + row) * 4) =
        make_float4(st_x0, st_sum, st_sq, 0.f);
  // look-ahead loads for the next tile, after this tile's last fence
  if constexpr (kLn) {
    lp_pending = ln_parts_next;
    if (ln_parts_next) ln_load_parts(p, ln_nrow, lp);
  }
  if constexpr (kPre) {
    if (next_m0 >= 0 && pre_base != nullptr) {  // the first half of the next tile's slice of this thread's row
      const int nrow = next_m0 + q * 32 + lane;
      if (nrow < p.M) {
        const uint4* r4 = reinterpret_cast<const uint4*>(pre_base + (size_t)nrow * pre_ld + next_n_base +
                                                         cg * kCols);
#pragma unroll
        for (int j = 0; j < kPreHalf; ++j) pre_lo[j] = r4[j];
      }
    }
  }
  ln_st = ln_next;
}

// Epilogue of the flavours that combine the accumulator with a SECOND global operand of the output's shape
// (kEpiResid: the residual stream; kEpiAct2: the saved pre-activation), GB_PRE_TMA build (default).
// Read by the thread that owns the row (16-byte pieces of 32 different rows per instruction, as the first
// version did) that operand cost the out-proj GEMM 23 % (B200, M = 94 700, N = K = 768: 95 µs plain / bias,
// 117 µs with the residual).  Here it arrives the way the output leaves: a 32 x 32 box per chunk, fetched by the
// TMA engine straight into the half-slab the chunk's output is then written to IN PLACE (each thread reads and
// rewrites its own 64 bytes) and stored from.  Five half-slabs per warp rotate; the box of chunk g + 3 is requested
// when chunk g starts — its slab was last read by the store of chunk g − 2, which wait_group.read<1> has seen off —
// so a box has three chunks (¾ of a tile) to arrive.  One mbarrier per half-slab (the lane that arms it is the
// lane that issues the load); no registers are spent on the operand.
constexpr int kPreSlabs = 5;
constexpr int kPreAhead = 3;
constexpr int kPreSlabBytes = 32 * 64;  // 32 rows x 32 fp16 columns

template <int kMode, typename WaitAcc, typename ReleaseAcc>
__device__ __forceinline__ void gemm_epilogue_tile_pre(const GemmParams& p, const CUtensorMap* tmC,
                                                       const CUtensorMap* tmR, uint8_t* slabs, uint64_t* pre_bar,
                                                       uint32_t tmem_acc, int m0, int n_base, int warp, int lane,
                                                       int next_m0, uint32_t consts_s, uint32_t consts_next_s,
                                                       int next_n_base, int& ring_s, uint32_t& ring_par,
                                                       WaitAcc&& wait_acc, ReleaseAcc&& release_acc) {
  static_assert(kMode == kEpiResid || kMode == kEpiAct2, "second-operand flavours only");
  constexpr int kCols = kEpiCols;
  constexpr int kChunks = kCols / 32;
  constexpr bool kAct2 = kMode == kEpiAct2;
  const int q = warp & 3;
  const int cg = (warp - 4) >> 2;
  const int n0 = n_base + cg * kCols;
  const int row0 = m0 + q * 32;
  const int row = row0 + lane;
  const bool row_ok = row < p.M;
  constexpr int kCPL = kCols / 32;
  float nb[kCPL];
#pragma unroll
  for (int i = 0; i < kCPL; ++i) nb[i] = 0.f;
  if (!kAct2 && next_n_base >= 0 && p.bias != nullptr) {
    const int nc = next_n_base + cg * kCols + kCPL * lane;
#pragma unroll
    for (int i = 0; i < kCPL; ++i) nb[i] = __ldg(p.bias + nc + i);
  }
  float st_sum = 0.f, st_sq = 0.f, st_x0 = 0.f;
  wait_acc();
  tc_fence_after();
  const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(q * 32) << 16) + cg * kCols;
  uint32_t v[kCols];
#pragma unroll
  for (int i = 0; i < kCols / 64; ++i) tmem_ld_32x64(taddr + 64 * i, v + 64 * i);
  tmem_ld_wait();
  tc_fence_before();
  __syncwarp();
  release_acc();
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
    const int col0 = n0 + c * 32;
    const int s = ring_s;
    if (lane == 0) {
      // the box of the chunk kPreAhead ahead → the slab the store of two chunks ago has finished reading
      tma_store_wait_read<1>();
      int sa = s + kPreAhead;
      if (sa >= kPreSlabs) sa -= kPreSlabs;
      const bool same_tile = c + kPreAhead < kChunks;
      const int la_row = same_tile ? row0 : next_m0 + q * 32;
      const int la_col = same_tile ? col0 + kPreAhead * 32 : next_n_base + cg * kCols + (c + kPreAhead - kChunks) * 32;
      if (same_tile || next_m0 >= 0) {
        mbar_expect_tx(&pre_bar[sa], kPreSlabBytes);
        tma_load_2d(slabs + sa * kPreSlabBytes, tmR, &pre_bar[sa], la_col, la_row);
      }
    }
    uint8_t* slab = slabs + s * kPreSlabBytes;
    const uint32_t slab_s = smem_u32(slab);
    mbar_wait(&pre_bar[s], (ring_par >> s) & 1u);
    ring_par ^= 1u << s;
    ring_s = s + 1 == kPreSlabs ? 0 : s + 1;
    uint4 pre[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) pre[j] = lds128u(slab_s + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4));
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[32 * c + j]);
    if (!kAct2 && p.bias != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 bb = lds128f(consts_s + (c * 32 + 4 * j) * 4);
        f[4 * j + 0] += bb.x; f[4 * j + 1] += bb.y; f[4 * j + 2] += bb.z; f[4 * j + 3] += bb.w;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __half2* h = reinterpret_cast<const __half2*>(&pre[j]);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 x = __half22float2(h[t]);
        if constexpr (kAct2) {
          f[8 * j + 2 * t] *= quick_gelu_grad(x.x);
          f[8 * j + 2 * t + 1] *= quick_gelu_grad(x.y);
        } else {
          f[8 * j + 2 * t] += x.x;
          f[8 * j + 2 * t + 1] += x.y;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 o;
      __half2* h = reinterpret_cast<__half2*>(&o);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        h[t] = __floats2half2_rn(f[8 * j + 2 * t], f[8 * j + 2 * t + 1]);
        if (kMode == kEpiResid && p.stats_out != nullptr) {  // statistics of the values as stored (fp16-rounded)
          const float2 r = __half22float2(h[t]);
          if (c == 0 && j == 0 && t == 0) st_x0 = r.x;
          const float d0 = r.x - st_x0, d1 = r.y - st_x0;
          st_sum += d0 + d1;
          st_sq = fmaf(d0, d0, fmaf(d1, d1, st_sq));
        }
      }
      sts128(slab_s + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4), o);
    }
    fence_proxy_async();  // generic-proxy writes → visible to the TMA engine
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(tmC, slab, col0, row0);
      tma_store_commit();
    }
  }
  if (!kAct2 && next_n_base >= 0) {  // park the next tile's biases (all reads of this tile's are done)
#pragma unroll
    for (int i = 0; i < kCPL; ++i) sts32f(consts_next_s + (kCPL * lane + i) * 4, nb[i]);
    __syncwarp();
  }
  if (kMode == kEpiResid && p.stats_out != nullptr && row_ok)
    *reinterpret_cast<float4*>(p.stats_out + ((size_t)(n0 / kCols) * p.M + row) * 4) =
        make_float4(st_x0, st_sum, st_sq, 0.f);
}

// -------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2), N % 256 == 0: a CTA pair computes 256 x 256 output tiles.
// CTA h of the pair stages rows [m0 + 128·h, +128) of A and rows [n0 + 128·h, +128) of W per
// k-block (32 KB per stage instead of 48 KB), both signalling the LEADER's full barrier;
// the leader's single MMA thread issues 256x256x16 tcgen05.mma.cta_group::2 instructions whose
// accumulator rows live in each CTA's own TMEM; commits are multicast to both CTAs' barriers; each
// CTA's 8 epilogue warps drain their own 128 rows and release the accumulator on the leader.
//
// kPairs = 2 (cluster of 4): two pairs work on vertically adjacent 256-row tiles of the SAME 256
// W rows.  At the tensor pipe's rate a pair-only kernel pulls 64 B/clk/SM out of L2 (9.5 KB/clk over
// 148 SMs, above what the L2 slices deliver), which capped the K=768 GEMMs at 60-70 % tensor
// activity.  Here CTA (pair p, half h) fetches only 64 of its 128 W rows and TMA-multicasts them to
// CTA h of both pairs, so a CTA requests 24 KB instead of 32 KB per k-block.  A smem slot is then
// written by both pairs' producers, hence a slot is free only when BOTH leaders' MMAs have retired
// (empty barriers count kPairs arrivals; commits are multicast to the whole cluster).
// Per-row arithmetic is identical for kPairs = 1 and 2 (same k order, same epilogue).
// -------------------------------------------------------------------------------------------------
// (row block, column tile) of a persistent CTA's tile sequence tile = first, first + step, …, advanced
// incrementally: one integer division pair at start-up instead of two per tile on the critical path.
struct TileWalk {
  int mi, ni, dm, dn, n_tiles;
  __device__ __forceinline__ TileWalk(int first, int step, int n_tiles_) : n_tiles(n_tiles_) {
    mi = first / n_tiles; ni = first % n_tiles;
    dm = step / n_tiles; dn = step % n_tiles;
  }
  __device__ __forceinline__ void next() {
    mi += dm; ni += dn;
    if (ni >= n_tiles) { ni -= n_tiles; ++mi; }
  }
};

#ifndef GB_PRE_TMA
#define GB_PRE_TMA 1  // residual / saved pre-activation through TMA into the output slabs (gemm_epilogue_tile_pre)
#endif
template <int kMode>
struct Gemm2Cfg {
  static constexpr int BN = 256;
  static constexpr bool kPreTma = GB_PRE_TMA && (kMode == kEpiResid || kMode == kEpiAct2);
#ifndef GB_STAGES2
#define GB_STAGES2 ((GB_EPI_WARPS == 8 && GB_SLAB_COLS == 32) ? 5 : 4)  // measured: 4, 5 and 6 stages perform alike; 16 warps need the smem
#endif
  static constexpr int kStages = kPreTma ? 4 : GB_STAGES2;  // the five rotating half-slabs take a stage's room
  static constexpr int kSlabBytes = kPreTma ? kEpiWarps2 * kPreSlabs * kPreSlabBytes
                                            : kEpiWarps2 * 2 * 64 * GB_SLAB_COLS;  // two 32-row output (half-)slabs per epilogue warp
  static constexpr int kConstBytes = kEpiWarps2 * 2 * kEpiCols * 8;  // per warp, double-buffered: biases + column sums
  static constexpr int kPreBarBytes = kPreTma ? 512 : 0;             // one mbarrier per rotating half-slab
  static constexpr int kABytes = kBM * kBK * 2;        // 16 KB: this CTA's 128 rows of A
  static constexpr int kBBytes = (BN / 2) * kBK * 2;   // 16 KB: this CTA's half of the W tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + kSlabBytes + kConstBytes + 1024 + 256 + kPreBarBytes;
  static_assert(kSmemBytes <= 232448, "shared memory per CTA");
  static_assert(!kPreTma || kEpiWarps2 * kPreSlabs * 8 <= kPreBarBytes, "mbarrier room");
};

template <int kPairs, int kMode>
__global__ void __launch_bounds__(kGemm2Threads, 1)
gemm_f16_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tmA,
                             const __grid_constant__ CUtensorMap tmB,
                             const __grid_constant__ CUtensorMap tmC,
                             const __grid_constant__ CUtensorMap tmR, const GemmParams p) {
  using Cfg = Gemm2Cfg<kMode>;
  constexpr int kStages = Cfg::kStages;
  constexpr int BN = Cfg::BN;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * Cfg::kABytes;
  uint8_t* smem_slabs = smem + kStages * Cfg::kStageBytes;  // 1024-aligned: stage bytes are 32 KB
  uint8_t* smem_consts = smem_slabs + Cfg::kSlabBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_consts + Cfg::kConstBytes);
  uint64_t* full_bar = bars;                      // [kStages]  (used on the leader)
  uint64_t* empty_bar = bars + kStages;           // [kStages]  (each CTA waits on its own)
  uint64_t* tfull_bar = bars + 2 * kStages;       // [2]        (each CTA waits on its own)
  uint64_t* tempty_bar = bars + 2 * kStages + 2;  // [2]        (used on the leader)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
  uint64_t* pre_bars = reinterpret_cast<uint64_t*>(smem_consts + Cfg::kConstBytes + 256);  // [warps][kPreSlabs] (kPreTma)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t pair = rank >> 1;   // which 256-row tile of the cluster's row block
  const uint32_t half = rank & 1;    // which CTA of the MMA pair
  const bool leader = half == 0;
  constexpr int kClusterRows = 2 * kBM * kPairs;
  constexpr uint16_t kAllCtas = (1u << (2 * kPairs)) - 1;

  const int m_blocks = (p.M + kClusterRows - 1) / kClusterRows;
  const int n_tiles = p.N / BN;
  const int num_tiles = m_blocks * n_tiles;
  const int k_blocks = p.K / kBK;
  const int cluster_id = blockIdx.x / (2 * kPairs);
  const int num_clusters = gridDim.x / (2 * kPairs);
  const int row_off = pair * 2 * kBM + half * kBM;  // this CTA's rows inside the cluster's row block

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    if constexpr (Cfg::kPreTma) tma_prefetch_desc(&tmR);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);   // leader's arrive.expect_tx covering both CTAs' loads
      mbar_init(&empty_bar[s], kPairs);  // every leader's multicast commit
    }
    if constexpr (Cfg::kPreTma)
      for (int s = 0; s < kEpiWarps2 * kPreSlabs; ++s) mbar_init(&pre_bars[s], 1);  // the arming lane's arrive.expect_tx
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);    // the leader's multicast commit
      mbar_init(&tempty_bar[s], 2 * kEpiWarps2);  // the epilogue warps of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_2cta(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();  // barriers of BOTH CTAs are initialised before anything can signal them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: barrier set-up, TMEM allocation and descriptor prefetch above ran while the
  // previous kernel of the stream was still draining its last tiles; its results (A, residual, statistics)
  // are touched only from here on, and the next kernel may be scheduled as soon as SMs free up.
  pdl_wait();
  pdl_launch_dependents();

  if (warp < 4) reg_dealloc<kCtlRegs>();  // warpgroup 0 (TMA, MMA, TMEM-alloc, idle) hands its registers over
  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      GB_STALL_DECL(w_empty);
      TileWalk tw(cluster_id, num_clusters, n_tiles);
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, tw.next()) {
        const int m0 = tw.mi * kClusterRows + row_off;
        const int n0 = tw.ni * BN + half * (BN / 2);
        for (int kb = 0; kb < k_blocks; ++kb) {
          GB_STALL_T(t_e);
          mbar_wait(&empty_bar[stage], phase ^ 1);
          GB_STALL_ADD(w_empty, t_e);
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
          tma_load_2d_2cta(smem_a + stage * Cfg::kABytes, &tmA, &full_bar[stage], kb * kBK, m0);
          if constexpr (kPairs == 1) {
            tma_load_2d_2cta(smem_b + stage * Cfg::kBBytes, &tmB, &full_bar[stage], kb * kBK, n0);
          } else {
            // my 64-row slice of this half's W rows, delivered to CTA `half` of every pair
            constexpr int kSlice = (BN / 2) / kPairs;
            tma_load_2d_2cta_mc(smem_b + stage * Cfg::kBBytes + pair * (kSlice * kBK * 2), &tmB,
                                &full_bar[stage], kb * kBK, n0 + pair * kSlice,
                                static_cast<uint16_t>(0x5u << half));
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
      (void)0;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(2 * kBM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      GB_STALL_DECL(w_full);
      GB_STALL_DECL(w_tempty);
      GB_STALL_T(t_mma0);
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        GB_STALL_T(t_te);
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        GB_STALL_ADD(w_tempty, t_te);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          GB_STALL_T(t_f);
          mbar_wait(&full_bar[stage], phase);
          GB_STALL_ADD(w_full, t_f);
          tc_fence_after();
          const uint64_t adesc = umma_desc_k_sw128(smem_u32(smem_a + stage * Cfg::kABytes));
          const uint64_t bdesc = umma_desc_k_sw128(smem_u32(smem_b + stage * Cfg::kBBytes));
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k)
            umma_f16_2cta(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          umma_commit_2cta(&empty_bar[stage], kAllCtas);  // one of kPairs arrivals in every CTA
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit_2cta(&tfull_bar[as], static_cast<uint16_t>(0b11u << (2 * pair)));  // both CTAs of the pair
      }
#ifdef GB_GEMM_STALLS
      GB_STALL_PUT(0, w_full);
      GB_STALL_PUT(1, w_tempty);
      GB_STALL_PUT(2, clock64() - t_mma0);
      GB_STALL_PUT(7, 1);
#endif
    }
  } else if (warp >= 4) {
    // ===================== epilogue (8 warps, both CTAs; fp16 output through TMA stores) =====================
    reg_alloc<kEpiRegs>();
    int it = 0;
    GB_STALL_DECL(w_tfull);
    GB_STALL_T(t_epi0);
    float2 ln_st = make_float2(0.f, 1.f);
    if constexpr (kMode == kEpiLn || kMode == kEpiLnGelu) {
      const int r = (cluster_id / n_tiles) * kClusterRows + row_off + (warp & 3) * 32 + lane;
      if (cluster_id < num_tiles && r < p.M) {
        if (p.ln_parts != nullptr) {
          float4 lp[kLnMaxParts];
          ln_load_parts(p, r, lp);
          ln_st = ln_merge_parts(lp, p.ln_nparts, p.ln_seg_n);
        } else {
          ln_st = *reinterpret_cast<const float2*>(p.ln_stats + (size_t)r * 2);
        }
      }
    }
    constexpr int kStrip = kEpiCols * 8;  // one buffer of the warp's constants: biases | column sums
    const uint32_t consts_s = smem_u32(smem_consts + (warp - 4) * 2 * kStrip);
    if (kMode != kEpiAct2 && cluster_id < num_tiles) {  // the first tile's per-column constants
      constexpr int kCPL = kEpiCols / 32;
      const int nc = (cluster_id % n_tiles) * BN + ((warp - 4) >> 2) * kEpiCols + kCPL * lane;
#pragma unroll
      for (int i = 0; i < kCPL; ++i) {
        sts32f(consts_s + (kCPL * lane + i) * 4, p.bias != nullptr ? __ldg(p.bias + nc + i) : 0.f);
        if (kMode == kEpiLn || kMode == kEpiLnGelu)
          sts32f(consts_s + kEpiCols * 4 + (kCPL * lane + i) * 4, __ldg(p.col_sum + nc + i));
      }
      __syncwarp();
    }
    if constexpr (Cfg::kPreTma) {
      uint8_t* my_slabs = smem_slabs + (warp - 4) * (kPreSlabs * kPreSlabBytes);
      uint64_t* my_bars = pre_bars + (warp - 4) * kPreSlabs;
      int ring_s = 0;
      uint32_t ring_par = 0;
      if (lane == 0 && cluster_id < num_tiles) {  // the first kPreAhead boxes of the first tile
        const int r0 = (cluster_id / n_tiles) * kClusterRows + row_off + (warp & 3) * 32;
        const int c0 = (cluster_id % n_tiles) * BN + ((warp - 4) >> 2) * kEpiCols;
#pragma unroll
        for (int g = 0; g < kPreAhead; ++g) {
          mbar_expect_tx(&my_bars[g], kPreSlabBytes);
          tma_load_2d(my_slabs + g * kPreSlabBytes, &tmR, &my_bars[g], c0 + 32 * g, r0);
        }
      }
      TileWalk tw(cluster_id, num_clusters, n_tiles);
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        const int m0 = tw.mi * kClusterRows + row_off;
        const int n0 = tw.ni * BN;
        tw.next();
        const bool more = tile + num_clusters < num_tiles;
        const int next_m0 = more ? tw.mi * kClusterRows + row_off : -1;
        const int next_n0 = more ? tw.ni * BN : -1;
        gemm_epilogue_tile_pre<kMode>(
            p, &tmC, &tmR, my_slabs, my_bars, tmem_base + as * BN, m0, n0, warp, lane, next_m0,
            consts_s + (it & 1) * kStrip, consts_s + ((it + 1) & 1) * kStrip, next_n0, ring_s, ring_par,
            [&]() {
              GB_STALL_T(t_tf);
              mbar_wait(&tfull_bar[as], aphase);
              GB_STALL_ADD(w_tfull, t_tf);
            },
            [&]() {
              if (lane == 0) mbar_arrive_cluster(&tempty_bar[as], rank & ~1u);
            });
      }
    } else {
      float4 lp[kLnMaxParts];
      bool lp_pending = false;
      uint4 pre_lo[kEpiCols / 16];
      if ((kMode == kEpiResid || kMode == kEpiAct2) && cluster_id < num_tiles) {  // first tile's columns 0-63
        const __half* pre_base = kMode == kEpiAct2 ? p.aux : p.resid;
        const int pre_ld = kMode == kEpiAct2 ? p.ldo : p.ldr;
        const int r = (cluster_id / n_tiles) * kClusterRows + row_off + (warp & 3) * 32 + lane;
        if (pre_base != nullptr && r < p.M) {
          const uint4* r4 = reinterpret_cast<const uint4*>(
              pre_base + (size_t)r * pre_ld + (cluster_id % n_tiles) * BN + ((warp - 4) >> 2) * kEpiCols);
#pragma unroll
          for (int j = 0; j < kEpiCols / 16; ++j) pre_lo[j] = r4[j];
        }
      }
      TileWalk tw(cluster_id, num_clusters, n_tiles);
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        const int m0 = tw.mi * kClusterRows + row_off;
        const int n0 = tw.ni * BN;
        tw.next();
        const bool more = tile + num_clusters < num_tiles;
        const int next_m0 = more ? tw.mi * kClusterRows + row_off : -1;
        const int next_n0 = more ? tw.ni * BN : -1;
        gemm_epilogue_tile_tma<kMode>(
            p, &tmC, smem_slabs + (warp - 4) * (128 * GB_SLAB_COLS), tmem_base + as * BN, m0, n0, warp, lane, ln_st, next_m0,
            consts_s + (it & 1) * kStrip, consts_s + ((it + 1) & 1) * kStrip, next_n0, pre_lo, lp, lp_pending,
            [&]() {
              GB_STALL_T(t_tf);
              mbar_wait(&tfull_bar[as], aphase);
              GB_STALL_ADD(w_tfull, t_tf);
            },
            [&]() {
              if (lane == 0) mbar_arrive_cluster(&tempty_bar[as], rank & ~1u);
            });
      }
    }
#ifdef GB_GEMM_STALLS
    if (leader && warp == 4 && lane == 0) {
      GB_STALL_PUT(4, w_tfull);
      GB_STALL_PUT(5, clock64() - t_epi0);
    }
#endif
    if (lane == 0) tma_store_wait_all<0>();  // all output tiles are in global memory before exit
  }

  tc_fence_before();
  cluster_sync_all();  // nobody leaves while the peer may still read its smem or signal its barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace gb
