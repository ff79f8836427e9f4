// Host launcher + C-ABI entry for the tcgen05 GEMM.
#include <stdlib.h>
#include "ctx.h"
#include "gemm_tcgen05.cuh"

using namespace gb;

// Programmatic dependent launch of the GEMMs (GB_PDL=0 turns it off, for A/B measurements).
static bool gb_pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("GB_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}


template <int BN>
static int launch_gemm_bn(gb_ctx* c, const CUtensorMap& tmA, const CUtensorMap& tmB,
                          const GemmParams& p, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set[16] = {false};
  if (!attr_set[c->device & 15]) {
    GB_CUDA(c, cudaFuncSetAttribute(gemm_f16_tcgen05_kernel<BN>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set[c->device & 15] = true;
  }
  const int m_tiles = (p.M + kBM - 1) / kBM;
  const int tiles = m_tiles * (p.N / BN);
  const int sms = gb_gemm_sms(c);
  const int grid = tiles < sms ? tiles : sms;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes; cfg.stream = st;
  cfg.attrs = attr; cfg.numAttrs = gb_pdl_enabled() ? 1 : 0;
  {
    gb_prof_scope prof(c, st, 0, 2.0 * p.M * p.N * p.K, p.M, p.N, p.K);
    GB_CUDA(c, cudaLaunchKernelEx(&cfg, gemm_f16_tcgen05_kernel<BN>, tmA, tmB, p));
  }
  GB_LAUNCH_CHECK(c);
  return GB_OK;
}

template <int kMode>
static int launch_gemm_2cta(gb_ctx* c, const CUtensorMap& tmA, const CUtensorMap& tmB,
                            const CUtensorMap& tmC_in, const GemmParams& p, cudaStream_t st) {
  using Cfg = Gemm2Cfg<kMode>;
  // The flavours with a second operand of the output's shape (residual, saved pre-activation) fetch it through
  // a tensor map of their own, in the 32 x 32 boxes their outputs are stored in (gemm_epilogue_tile_pre).
  CUtensorMap tmC = tmC_in, tmR = tmC_in;
  if (Cfg::kPreTma) {
    const void* pre = kMode == kEpiAct2 ? (const void*)p.aux : (const void*)p.resid;
    const int pre_ld = kMode == kEpiAct2 ? p.ldo : p.ldr;
    if (!pre) return gb_fail(c, GB_ERR_ARG, "gemm: the second operand of this epilogue is missing");
    int rc = gb_make_tmap_2d_f16(c, &tmC, p.out, (uint64_t)p.M, (uint64_t)p.N, (uint64_t)p.ldo, 32, 32);
    if (rc) return rc;
    rc = gb_make_tmap_2d_f16(c, &tmR, pre, (uint64_t)p.M, (uint64_t)p.N, (uint64_t)pre_ld, 32, 32);
    if (rc) return rc;
  }
  // One pair per cluster.  (The kernel also supports two pairs sharing their W tile through TMA
  // multicast; measured on B200 at M = 51200, K = 768 / 3072 it gains nothing — the pair kernel is not
  // bound by L2 bandwidth — and four-CTA clusters strand 16 of the 148 SMs.)
  constexpr int kPairs = 1;
  constexpr int kCluster = 2 * kPairs;
  auto kernel = gemm_f16_tcgen05_2cta_kernel<kPairs, kMode>;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // see pdl_wait() in the kernel
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.blockDim = dim3(kGemm2Threads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = st;
  cfg.attrs = attr; cfg.numAttrs = gb_pdl_enabled() ? 2 : 1;
  static bool attr_set[16] = {false};
  if (!attr_set[c->device & 15]) {
    GB_CUDA(c, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set[c->device & 15] = true;
  }
  const int m_blocks = (p.M + 2 * kBM * kPairs - 1) / (2 * kBM * kPairs);
  const int tiles = m_blocks * (p.N / Cfg::BN);
  const int max_clusters = gb_gemm_sms(c) / kCluster;
  const int clusters = tiles < max_clusters ? tiles : max_clusters;
  cfg.gridDim = dim3(kCluster * clusters);
  {
    gb_prof_scope prof(c, st, 0, 2.0 * p.M * p.N * p.K, p.M, p.N, p.K);
    GB_CUDA(c, cudaLaunchKernelEx(&cfg, kernel, tmA, tmB, tmC, tmR, p));
  }
  GB_LAUNCH_CHECK(c);
  return GB_OK;
}

int gb_launch_gemm(gb_ctx* c, const void* A, int lda, const void* W, int ldw, const float* bias,
                   const void* resid, int ldr, void* out, int ldo, int M, int N, int K, int act,
                   int out_f32, cudaStream_t st, void* aux, const gb_gemm_ln* ln) {
  if (!A || !W || !out) return gb_fail(c, GB_ERR_ARG, "gemm: null pointer");
  if (M <= 0) return GB_OK;
  if (K % kBK != 0 || N % 128 != 0 || lda % 8 != 0 || ldw % 8 != 0 || ldo % 8 != 0 ||
      (resid && ldr % 8 != 0))
    return gb_fail(c, GB_ERR_ARG, "gemm: unsupported shape M=%d N=%d K=%d lda=%d ldw=%d ldo=%d", M,
                   N, K, lda, ldw, ldo);
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W) |
       reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(resid) |
       reinterpret_cast<uintptr_t>(bias)) & 15)
    return gb_fail(c, GB_ERR_ARG, "gemm: pointers must be 16-byte aligned");
  // 256-wide tiles halve the per-FLOP shared-memory traffic.  The tile shape is a function of N only:
  // a row's arithmetic must not depend on how many other rows are in the batch (pseudolabels have to
  // be bit-identical however the pool is batched or sharded across GPUs).
  // (fp32 outputs — the final feature projections, a few MFLOP — use the single-CTA kernel's direct-store
  // epilogue; the CTA-pair kernel only has the fp16 TMA-store epilogue)
  const bool wide = (N % 256 == 0) && !out_f32;
  // wide tiles run on CTA pairs (each CTA stages half of the 256 W rows); otherwise one CTA, BN = 128
  const int BN = 128;
  CUtensorMap tmA, tmB;
  int rc = gb_make_tmap_2d_f16(c, &tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, kBM);
  if (rc) return rc;
  rc = gb_make_tmap_2d_f16(c, &tmB, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, BN);
  if (rc) return rc;
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.out = out; p.ldo = ldo;
  p.bias = bias;
  p.resid = reinterpret_cast<const __half*>(resid); p.ldr = ldr;
  p.act = act; p.out_f32 = out_f32;
  p.aux = reinterpret_cast<__half*>(aux);
  p.ln_stats = nullptr; p.col_sum = nullptr; p.stats_out = nullptr;
  p.ln_parts = nullptr; p.ln_nparts = 0; p.ln_seg_n = 1.f;
  if (ln) {
    if (out_f32 || N % 256 != 0)
      return gb_fail(c, GB_ERR_ARG, "gemm: LayerNorm folding needs fp16 output and N %% 256 == 0");
    const bool fold_in = ln->ln_stats != nullptr || ln->ln_parts != nullptr;
    if (fold_in != (ln->col_sum != nullptr) || (ln->ln_stats && ln->ln_parts) ||
        (ln->ln_parts && (ln->nparts < 1 || ln->nparts > 768 / GB_STAT_SEG || K % ln->nparts != 0)))
      return gb_fail(c, GB_ERR_ARG, "gemm: inconsistent LayerNorm folding arguments");
    p.ln_stats = ln->ln_stats; p.col_sum = ln->col_sum;
    if (ln->ln_parts) {
      p.ln_parts = reinterpret_cast<const float4*>(ln->ln_parts);
      p.ln_nparts = ln->nparts;
      p.ln_seg_n = (float)(K / ln->nparts);
    }
    p.stats_out = ln->stats_out;
  }
  if (act == 2 && !aux) return gb_fail(c, GB_ERR_ARG, "gemm: act 2 needs aux");
  if (aux && (out_f32 || (reinterpret_cast<uintptr_t>(aux) & 15))) return gb_fail(c, GB_ERR_ARG, "gemm: aux needs fp16 output layout and 16-byte alignment");
  if (!wide) return launch_gemm_bn<128>(c, tmA, tmB, p, st);
  // output tensor map for the TMA-store epilogue: 32-column x 32-row boxes (64 B swizzle)
  CUtensorMap tmC;
  rc = gb_make_tmap_2d_f16(c, &tmC, out, (uint64_t)M, (uint64_t)N, (uint64_t)ldo, 32, GB_SLAB_COLS);
  if (rc) return rc;
  const bool fold = p.ln_stats != nullptr || p.ln_parts != nullptr;
  if (p.stats_out != nullptr && (fold || act != 0 || !resid))
    return gb_fail(c, GB_ERR_ARG, "gemm: row statistics are emitted by the residual epilogue only");
  if (act == 2) {
    if (fold || resid || bias) return gb_fail(c, GB_ERR_ARG, "gemm: act 2 takes no bias / residual / LayerNorm");
    return launch_gemm_2cta<kEpiAct2>(c, tmA, tmB, tmC, p, st);
  }
  if (act == 1) {
    if (resid) return gb_fail(c, GB_ERR_ARG, "gemm: act 1 with a residual is not supported");
    return fold ? launch_gemm_2cta<kEpiLnGelu>(c, tmA, tmB, tmC, p, st)
                : launch_gemm_2cta<kEpiGelu>(c, tmA, tmB, tmC, p, st);
  }
  if (resid) {
    if (fold) return gb_fail(c, GB_ERR_ARG, "gemm: folded LayerNorm with a residual is not supported");
    return launch_gemm_2cta<kEpiResid>(c, tmA, tmB, tmC, p, st);
  }
  return fold ? launch_gemm_2cta<kEpiLn>(c, tmA, tmB, tmC, p, st)
              : launch_gemm_2cta<kEpiPlain>(c, tmA, tmB, tmC, p, st);
}

// Debug hook (meaningful only in a -DGB_GEMM_STALLS build): accumulated stall counters of the CTA-pair
// kernel since the last reset: [0] MMA waits for operands, [1] MMA waits for a drained accumulator,
// [2] MMA thread total, [3] producer waits for a free slot, [4] epilogue warp 4 waits for the accumulator,
// [5] epilogue warp 4 total, [7] number of MMA threads that reported.
extern "C" int gb_debug_gemm_stalls(unsigned long long* out16, int reset) {
#ifdef GB_GEMM_STALLS
  if (cudaDeviceSynchronize() != cudaSuccess) return GB_ERR_CUDA;
  if (out16 && cudaMemcpyFromSymbol(out16, g_gemm_stalls, 128) != cudaSuccess) return GB_ERR_CUDA;
  if (reset) {
    unsigned long long z[16] = {0};
    if (cudaMemcpyToSymbol(g_gemm_stalls, z, 128) != cudaSuccess) return GB_ERR_CUDA;
  }
  return GB_OK;
#else
  (void)out16; (void)reset;
  return GB_ERR_ARG;
#endif
}

extern "C" int gb_gemm_f16(gb_ctx* c, const void* A, int lda, const void* W, int ldw,
                           const float* bias, const void* resid, int ldr, void* out, int ldo, int M,
                           int N, int K, int act, int out_f32, void* stream) {
  if (!c) return GB_ERR_ARG;
  gb_dev_guard dev_guard(c);
  if (act < 0 || act > 1) return gb_fail(c, GB_ERR_ARG, "gemm: act must be 0 or 1");
  return gb_launch_gemm(c, A, lda, W, ldw, bias, resid, ldr, out, ldo, M, N, K, act, out_f32,
                        reinterpret_cast<cudaStream_t>(stream), nullptr);
}
