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
    // TMEM lane quadrant q = warp % 4 (hardware rule); the two warps of a quadrant split the BN
    // columns.  The operand that has to come from global memory (residual, or the saved
    // pre-activation for act 2) is prefetched one 32-column chunk ahead — the first chunk before the
    // accumulator is even ready — so its latency hides behind the TMEM read-out of the previous one.
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    constexpr int kChunks = BN / 64;  // 32-column chunks per warp
    const __half* pre_base = (p.act == 2) ? p.aux : p.resid;
    const int pre_ld = (p.act == 2) ? p.ldo : p.ldr;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int m0 = (tile / n_tiles) * kBM;
      const int n0 = (tile % n_tiles) * BN + half * (BN / 2);
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      const bool has_pre = pre_base != nullptr && row_ok;
      uint4 pre[4], pre_next[4];
      if (has_pre) {
        const uint4* r4 = reinterpret_cast<const uint4*>(pre_base + (size_t)row * pre_ld + n0);
#pragma unroll
        for (int j = 0; j < 4; ++j) pre[j] = r4[j];
      }
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr =
          tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + half * (BN / 2);
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

}  // namespace gb
