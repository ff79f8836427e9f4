// Multi-head self-attention core for the short CLIP sequences (vision L = 50+P ≤ 96, text L ≤ 77,
// head dim 64):   O = softmax(Q·Kᵀ / 8 [+ causal mask]) · V    per (sample, head).
//
// Reference: nn.MultiheadAttention inside clip.model.ResidualAttentionBlock (need_weights=False,
// text tower: additive −inf upper-triangular mask) reached from models/clip_encoders.py:75-84,186.
// ≈1 % of the tower FLOPs: one CTA per (sample, head) keeps Q, K and Vᵀ of the
// whole sequence in shared memory and runs both contractions on mma.sync m16n8k16 register
// fragments (a 128-row tcgen05 tile would be >50 % padding at these lengths); the probabilities
// never leave registers.  Input is the packed in-proj output qkv[B·L, 3·D]; output a[B·L, D].
#include "ctx.h"
#include "common.cuh"

using namespace gb;

namespace {

constexpr int kDh = 64;
constexpr int kQKld = kDh + 8;  // padded row (halves) → conflict-free 32-bit fragment loads

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float x, float y) {
  const __half2 h = __floats2half2_rn(x, y);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t lds32(const __half* p) {
  return *reinterpret_cast<const uint32_t*>(p);
}

// Stage Q, K, V (row-major, rows padded to 72 halves) of one (sample, head) in shared memory; rows ≥ L
// are zero.  The transposed operands the PV product needs come from ldmatrix.trans, not from a
// transposed copy (2-byte scattered stores were 8-way bank conflicted).
template <int NT>
__device__ __forceinline__ void stage_qkv(const __half* __restrict__ qkv, int L, int D, int b,
                                          int h, __half* Qs, __half* Ks, __half* Vs) {
  constexpr int Lp = NT * 8;
  const size_t ld = (size_t)3 * D;
  // all global loads of the tile are issued before the first shared-memory store (128 threads, one
  // 16-byte chunk of Q, K and V each per step): the tile arrives in one round trip instead of kIt
  constexpr int kIt = (Lp * 8) / 128;
  uint4 q[kIt], k[kIt], v[kIt];
#pragma unroll
  for (int it = 0; it < kIt; ++it) {
    const int i = it * 128 + threadIdx.x;
    const int r = i >> 3, ch = i & 7;
    q[it] = k[it] = v[it] = make_uint4(0, 0, 0, 0);
    if (r < L) {
      const __half* src = qkv + ((size_t)b * L + r) * ld + h * kDh + ch * 8;
      q[it] = *reinterpret_cast<const uint4*>(src);
      k[it] = *reinterpret_cast<const uint4*>(src + D);
      v[it] = *reinterpret_cast<const uint4*>(src + 2 * D);
    }
  }
#pragma unroll
  for (int it = 0; it < kIt; ++it) {
    const int i = it * 128 + threadIdx.x;
    const int r = i >> 3, ch = i & 7;
    *reinterpret_cast<uint4*>(Qs + r * kQKld + ch * 8) = q[it];
    *reinterpret_cast<uint4*>(Ks + r * kQKld + ch * 8) = k[it];
    *reinterpret_cast<uint4*>(Vs + r * kQKld + ch * 8) = v[it];
  }
}

// Four transposed 8x8 b16 tiles in one instruction; lane l supplies the row address of row l%8 of
// tile l/8.  With .trans, thread (g = lane/4, t = lane%4) receives {M[2t][g], M[2t+1][g]} of each
// tile — exactly the "col" B fragment of mma.m16n8k16 when M is stored [k][n] row-major.
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const __half* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}

// S = scale·Q·Kᵀ with masking, then row soft-max, all in the m16n8 accumulator layout:
// thread (g = lane/4, t = lane%4) owns rows r0+g (s[..][0,1]) and r0+g+8 (s[..][2,3]),
// columns nt*8 + 2t, +1.  Returns the row sums' inverses; s holds exp(x − max).
template <int NT>
__device__ __forceinline__ void scores_softmax(const __half* Qs, const __half* Ks, int r0, int L,
                                               bool causal, int lane, float (&s)[NT][4],
                                               float& inv0, float& inv1) {
  const int g = lane >> 2, t = lane & 3;
  uint32_t a[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    a[ks][0] = lds32(Qs + (r0 + g) * kQKld + ks * 16 + 2 * t);
    a[ks][1] = lds32(Qs + (r0 + g + 8) * kQKld + ks * 16 + 2 * t);
    a[ks][2] = lds32(Qs + (r0 + g) * kQKld + ks * 16 + 8 + 2 * t);
    a[ks][3] = lds32(Qs + (r0 + g + 8) * kQKld + ks * 16 + 8 + 2 * t);
  }
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const uint32_t b0 = lds32(Ks + (nt * 8 + g) * kQKld + ks * 16 + 2 * t);
      const uint32_t b1 = lds32(Ks + (nt * 8 + g) * kQKld + ks * 16 + 8 + 2 * t);
      mma16816(s[nt], a[ks], b0, b1);
    }
    const int c0 = nt * 8 + 2 * t;
    const int ra = r0 + g, rb = r0 + g + 8;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int col = c0 + e;
      const bool oka = col < L && !(causal && col > ra);
      const bool okb = col < L && !(causal && col > rb);
      // scores in log2 units: (q·k / 8)·log2(e), so the soft-max needs a bare ex2
      s[nt][e] = oka ? s[nt][e] * 0.18033688011112042f : -INFINITY;
      s[nt][2 + e] = okb ? s[nt][2 + e] * 0.18033688011112042f : -INFINITY;
      m0 = fmaxf(m0, s[nt][e]);
      m1 = fmaxf(m1, s[nt][2 + e]);
    }
  }
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
  float l0 = 0.f, l1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      s[nt][e] = fast_exp2(s[nt][e] - m0);       // col 0 is never masked → m finite
      s[nt][2 + e] = fast_exp2(s[nt][2 + e] - m1);
      l0 += s[nt][e];
      l1 += s[nt][2 + e];
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  inv0 = 1.0f / l0;
  inv1 = 1.0f / l1;
}

template <int NT>
__global__ void __launch_bounds__(128)
attn_fwd_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int L, int D,
                int causal) {
  pdl_launch_dependents();  // the out-proj GEMM behind us may set itself up while we drain
  constexpr int Lp = NT * 8;  // multiple of 16
  extern __shared__ __align__(16) uint8_t attn_smem[];
  __half* Qs = reinterpret_cast<__half*>(attn_smem);
  __half* Ks = Qs + Lp * kQKld;
  __half* Vs = Ks + Lp * kQKld;
  // Samples are visited from the LAST one down: the in-proj GEMM in front of this kernel writes its rows in
  // ascending order, so the rows it wrote last (still in the 126 MB L2) are the ones read first here — and this
  // kernel's own last writes (sample 0 …) are what the out-proj GEMM behind it reads first.
  const int h = blockIdx.x, b = gridDim.y - 1 - blockIdx.y;
  stage_qkv<NT>(qkv, L, D, b, h, Qs, Ks, Vs);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  for (int mt = warp; mt * 16 < L; mt += 4) {
    const int r0 = mt * 16;
    float s[NT][4];
    float inv0, inv1;
    scores_softmax<NT>(Qs, Ks, r0, L, causal != 0, lane, s, inv0, inv1);
    float o[8][4];
#pragma unroll
    for (int dn = 0; dn < 8; ++dn) o[dn][0] = o[dn][1] = o[dn][2] = o[dn][3] = 0.f;
#pragma unroll
    for (int kt = 0; kt < NT / 2; ++kt) {
      uint32_t a[4];
      a[0] = pack_h2(s[2 * kt][0] * inv0, s[2 * kt][1] * inv0);
      a[1] = pack_h2(s[2 * kt][2] * inv1, s[2 * kt][3] * inv1);
      a[2] = pack_h2(s[2 * kt + 1][0] * inv0, s[2 * kt + 1][1] * inv0);
      a[3] = pack_h2(s[2 * kt + 1][2] * inv1, s[2 * kt + 1][3] * inv1);
      // B[k = key][n = dh] = V[key][dh]: tiles (keys 0-7 | 8-15) × (dh block dn | dn+1)
      const __half* vrow = Vs + (kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * kQKld + (lane >> 4) * 8;
#pragma unroll
      for (int dn = 0; dn < 8; dn += 2) {
        uint32_t bf[4];
        ldmatrix_x4_trans(bf, vrow + dn * 8);
        mma16816(o[dn], a, bf[0], bf[1]);
        mma16816(o[dn + 1], a, bf[2], bf[3]);
      }
    }
    // stage the 16x64 output tile in this warp's own (no longer needed) Q rows, then write whole
    // 128-byte rows
    __syncwarp();
#pragma unroll
    for (int dn = 0; dn < 8; ++dn) {
      *reinterpret_cast<uint32_t*>(Qs + (r0 + g) * kQKld + dn * 8 + 2 * t) = pack_h2(o[dn][0], o[dn][1]);
      *reinterpret_cast<uint32_t*>(Qs + (r0 + g + 8) * kQKld + dn * 8 + 2 * t) = pack_h2(o[dn][2], o[dn][3]);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + i * 4 + (lane >> 3), ch = lane & 7;
      if (r < L)
        *reinterpret_cast<uint4*>(out + ((size_t)b * L + r) * D + h * kDh + ch * 8) =
            *reinterpret_cast<const uint4*>(Qs + r * kQKld + ch * 8);
    }
  }
}

// -------------------------------------------------------------------------------------------------
// Backward w.r.t. the packed qkv (frozen weights: only data gradients exist).
//   P = softmax(S),  dV = Pᵀ·dO,  dP = dO·Vᵀ,  dS = P ∘ (dP − rowsum(dP∘P)),  dQ = dS·K/8,  dK = dSᵀ·Q/8
// Same one-CTA-per-(sample, head) layout; P is recomputed from the saved qkv.  dS is staged through
// shared memory so the transposed products can re-use the m16n8k16 fragments.
// -------------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(128)
attn_bwd_kernel(const __half* __restrict__ qkv, const __half* __restrict__ dout,
                __half* __restrict__ dqkv, int L, int D, int causal) {
  pdl_launch_dependents();
  constexpr int Lp = NT * 8;
  constexpr int ldp = Lp + 8;   // row stride of the [Lp, Lp] P / dS tiles
  extern __shared__ __align__(16) uint8_t attn_smem[];
  __half* Qs = reinterpret_cast<__half*>(attn_smem);   // [Lp, 72] row-major, like K, V, dO
  __half* Ks = Qs + Lp * kQKld;
  __half* Vs = Ks + Lp * kQKld;
  __half* dOs = Vs + Lp * kQKld;
  __half* Ps = dOs + Lp * kQKld;                       // [Lp, ldp]  P  (query-major)
  __half* dSs = Ps + Lp * ldp;                         // [Lp, ldp]  dS (query-major)
  const int h = blockIdx.x, b = gridDim.y - 1 - blockIdx.y;   // last sample first (see attn_fwd_kernel)
  const size_t ld = (size_t)3 * D;
  {
    // one round trip: every 16-byte chunk of Q, K, V, dO is requested before the first store; all
    // transposed operands below come from ldmatrix.trans on these row-major tiles
    constexpr int kIt = (Lp * 8) / 128;
    uint4 q[kIt], k[kIt], v[kIt], d[kIt];
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
      const int i = it * 128 + threadIdx.x;
      const int r = i >> 3, ch = i & 7;
      q[it] = k[it] = v[it] = d[it] = make_uint4(0, 0, 0, 0);
      if (r < L) {
        const __half* src = qkv + ((size_t)b * L + r) * ld + h * kDh + ch * 8;
        q[it] = *reinterpret_cast<const uint4*>(src);
        k[it] = *reinterpret_cast<const uint4*>(src + D);
        v[it] = *reinterpret_cast<const uint4*>(src + 2 * D);
        d[it] = *reinterpret_cast<const uint4*>(dout + ((size_t)b * L + r) * D + h * kDh + ch * 8);
      }
    }
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
      const int i = it * 128 + threadIdx.x;
      const int r = i >> 3, ch = i & 7;
      *reinterpret_cast<uint4*>(Qs + r * kQKld + ch * 8) = q[it];
      *reinterpret_cast<uint4*>(Ks + r * kQKld + ch * 8) = k[it];
      *reinterpret_cast<uint4*>(Vs + r * kQKld + ch * 8) = v[it];
      *reinterpret_cast<uint4*>(dOs + r * kQKld + ch * 8) = d[it];
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  // lane-dependent parts of the ldmatrix.x4.trans row addresses
  const int lm_row = (lane & 7) + ((lane >> 3) & 1) * 8;   // B operand: k rows 0-7 | 8-15
  const int lm_col = (lane >> 4) * 8;                      //            n block dn | dn+1
  const int la_row = (lane & 7) + ((lane >> 4) & 1) * 8;   // transposed A operand: k (query) rows
  const int la_col = ((lane >> 3) & 1) * 8;                //                       m (key) block
  // ---- phase 1: per query-row tile: P, dP, dS (→ smem), dQ (→ global) ---------------------------
  for (int mt = warp; mt * 16 < Lp; mt += 4) {
    const int r0 = mt * 16;
    float s[NT][4];
    float inv0, inv1;
    scores_softmax<NT>(Qs, Ks, r0, L, causal != 0, lane, s, inv0, inv1);
    // dP = dO · Vᵀ   (A = dO rows, B[k=dh][n=key] = V[key][dh] → V row-major is already "col")
    uint32_t a[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      a[ks][0] = lds32(dOs + (r0 + g) * kQKld + ks * 16 + 2 * t);
      a[ks][1] = lds32(dOs + (r0 + g + 8) * kQKld + ks * 16 + 2 * t);
      a[ks][2] = lds32(dOs + (r0 + g) * kQKld + ks * 16 + 8 + 2 * t);
      a[ks][3] = lds32(dOs + (r0 + g + 8) * kQKld + ks * 16 + 8 + 2 * t);
    }
    float dsum0 = 0.f, dsum1 = 0.f;
    float dp[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t b0 = lds32(Vs + (nt * 8 + g) * kQKld + ks * 16 + 2 * t);
        const uint32_t b1 = lds32(Vs + (nt * 8 + g) * kQKld + ks * 16 + 8 + 2 * t);
        mma16816(dp[nt], a[ks], b0, b1);
      }
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        s[nt][e] *= inv0;          // P
        s[nt][2 + e] *= inv1;
        dsum0 += dp[nt][e] * s[nt][e];
        dsum1 += dp[nt][2 + e] * s[nt][2 + e];
      }
    }
    dsum0 += __shfl_xor_sync(0xffffffffu, dsum0, 1);
    dsum0 += __shfl_xor_sync(0xffffffffu, dsum0, 2);
    dsum1 += __shfl_xor_sync(0xffffffffu, dsum1, 1);
    dsum1 += __shfl_xor_sync(0xffffffffu, dsum1, 2);
    const bool va = r0 + g < L, vb = r0 + g + 8 < L;  // padded query rows contribute nothing
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int c0 = nt * 8 + 2 * t;
      const float p0 = va ? s[nt][0] : 0.f, p1 = va ? s[nt][1] : 0.f;
      const float p2 = vb ? s[nt][2] : 0.f, p3 = vb ? s[nt][3] : 0.f;
      const float d0 = p0 * (dp[nt][0] - dsum0), d1 = p1 * (dp[nt][1] - dsum0);
      const float d2 = p2 * (dp[nt][2] - dsum1), d3 = p3 * (dp[nt][3] - dsum1);
      *reinterpret_cast<uint32_t*>(Ps + (r0 + g) * ldp + c0) = pack_h2(p0, p1);
      *reinterpret_cast<uint32_t*>(Ps + (r0 + g + 8) * ldp + c0) = pack_h2(p2, p3);
      *reinterpret_cast<uint32_t*>(dSs + (r0 + g) * ldp + c0) = pack_h2(d0, d1);
      *reinterpret_cast<uint32_t*>(dSs + (r0 + g + 8) * ldp + c0) = pack_h2(d2, d3);
      s[nt][0] = d0; s[nt][1] = d1; s[nt][2] = d2; s[nt][3] = d3;
    }
    // dQ = dS · K / 8   (A = dS from registers, B[k=key][n=dh] = K[key][dh] → ldmatrix.trans on K)
    float o[8][4];
#pragma unroll
    for (int dn = 0; dn < 8; ++dn) o[dn][0] = o[dn][1] = o[dn][2] = o[dn][3] = 0.f;
#pragma unroll
    for (int kt = 0; kt < NT / 2; ++kt) {
      uint32_t af[4];
      af[0] = pack_h2(s[2 * kt][0], s[2 * kt][1]);
      af[1] = pack_h2(s[2 * kt][2], s[2 * kt][3]);
      af[2] = pack_h2(s[2 * kt + 1][0], s[2 * kt + 1][1]);
      af[3] = pack_h2(s[2 * kt + 1][2], s[2 * kt + 1][3]);
      const __half* krow = Ks + (kt * 16 + lm_row) * kQKld + lm_col;
#pragma unroll
      for (int dn = 0; dn < 8; dn += 2) {
        uint32_t bf[4];
        ldmatrix_x4_trans(bf, krow + dn * 8);
        mma16816(o[dn], af, bf[0], bf[1]);
        mma16816(o[dn + 1], af, bf[2], bf[3]);
      }
    }
#pragma unroll
    for (int dn = 0; dn < 8; ++dn) {
      const int col = h * kDh + dn * 8 + 2 * t;
      if (va)
        *reinterpret_cast<__half2*>(dqkv + ((size_t)b * L + r0 + g) * ld + col) =
            __floats2half2_rn(o[dn][0] * 0.125f, o[dn][1] * 0.125f);
      if (vb)
        *reinterpret_cast<__half2*>(dqkv + ((size_t)b * L + r0 + g + 8) * ld + col) =
            __floats2half2_rn(o[dn][2] * 0.125f, o[dn][3] * 0.125f);
    }
  }
  __syncthreads();
  // ---- phase 2: per key-row tile: dK = dSᵀ·Q/8, dV = Pᵀ·dO -------------------------------------
  // A[m=key][k=query] = dS[query][key] (resp. P): the transposed fragments come from
  // ldmatrix.x4.trans on the query-major tiles; B[k=query][n=dh] = Q / dO rows, also via .trans.
  for (int mt = warp; mt * 16 < L; mt += 4) {
    const int r0 = mt * 16;  // key rows
    float ok[8][4], ov[8][4];
#pragma unroll
    for (int dn = 0; dn < 8; ++dn) {
      ok[dn][0] = ok[dn][1] = ok[dn][2] = ok[dn][3] = 0.f;
      ov[dn][0] = ov[dn][1] = ov[dn][2] = ov[dn][3] = 0.f;
    }
#pragma unroll
    for (int kt = 0; kt < NT / 2; ++kt) {
      const int q0 = kt * 16;  // query block
      uint32_t as[4], ap[4];
      ldmatrix_x4_trans(as, dSs + (q0 + la_row) * ldp + r0 + la_col);
      ldmatrix_x4_trans(ap, Ps + (q0 + la_row) * ldp + r0 + la_col);
      const __half* qrow = Qs + (q0 + lm_row) * kQKld + lm_col;
      const __half* drow = dOs + (q0 + lm_row) * kQKld + lm_col;
#pragma unroll
      for (int dn = 0; dn < 8; dn += 2) {
        uint32_t bq[4], bd[4];
        ldmatrix_x4_trans(bq, qrow + dn * 8);
        ldmatrix_x4_trans(bd, drow + dn * 8);
        mma16816(ok[dn], as, bq[0], bq[1]);
        mma16816(ok[dn + 1], as, bq[2], bq[3]);
        mma16816(ov[dn], ap, bd[0], bd[1]);
        mma16816(ov[dn + 1], ap, bd[2], bd[3]);
      }
    }
    const int ra = r0 + g, rb = r0 + g + 8;
#pragma unroll
    for (int dn = 0; dn < 8; ++dn) {
      const int col = h * kDh + dn * 8 + 2 * t;
      if (ra < L) {
        *reinterpret_cast<__half2*>(dqkv + ((size_t)b * L + ra) * ld + D + col) =
            __floats2half2_rn(ok[dn][0] * 0.125f, ok[dn][1] * 0.125f);
        *reinterpret_cast<__half2*>(dqkv + ((size_t)b * L + ra) * ld + 2 * D + col) =
            __floats2half2_rn(ov[dn][0], ov[dn][1]);
      }
      if (rb < L) {
        *reinterpret_cast<__half2*>(dqkv + ((size_t)b * L + rb) * ld + D + col) =
            __floats2half2_rn(ok[dn][2] * 0.125f, ok[dn][3] * 0.125f);
        *reinterpret_cast<__half2*>(dqkv + ((size_t)b * L + rb) * ld + 2 * D + col) =
            __floats2half2_rn(ov[dn][2], ov[dn][3]);
      }
    }
  }
}

template <int NT>
size_t fwd_smem() {
  constexpr int Lp = NT * 8;
  return (size_t)(3 * Lp * kQKld) * 2;
}
template <int NT>
size_t bwd_smem() {
  constexpr int Lp = NT * 8;
  return (size_t)(4 * Lp * kQKld + 2 * Lp * (Lp + 8)) * 2;
}

template <int NT>
int launch_fwd(gb_ctx* c, const void* qkv, void* out, int B, int L, int D, int causal,
               cudaStream_t st) {
  const size_t smem = fwd_smem<NT>();
  if (smem > 48 * 1024) {
    static bool done[16] = {false};
    if (!done[c->device & 15]) {
      GB_CUDA(c, cudaFuncSetAttribute(attn_fwd_kernel<NT>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      done[c->device & 15] = true;
    }
  }
  attn_fwd_kernel<NT><<<dim3(D / kDh, B), 128, smem, st>>>((const __half*)qkv, (__half*)out, L, D,
                                                          causal);
  GB_LAUNCH_CHECK(c);
  return GB_OK;
}
template <int NT>
int launch_bwd(gb_ctx* c, const void* qkv, const void* dout, void* dqkv, int B, int L, int D,
               int causal, cudaStream_t st) {
  const size_t smem = bwd_smem<NT>();
  if (smem > 48 * 1024) {
    static bool done[16] = {false};
    if (!done[c->device & 15]) {
      GB_CUDA(c, cudaFuncSetAttribute(attn_bwd_kernel<NT>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      done[c->device & 15] = true;
    }
  }
  attn_bwd_kernel<NT><<<dim3(D / kDh, B), 128, smem, st>>>((const __half*)qkv, (const __half*)dout,
                                                          (__half*)dqkv, L, D, causal);
  GB_LAUNCH_CHECK(c);
  return GB_OK;
}

}  // namespace

// attn_tc.cu: the tcgen05 / TMEM / TMA forward (default); GB_ATTN_LEGACY=1 keeps the mma.sync kernel below
bool gb_attn_tc_enabled();
int gb_launch_attn_fwd_tc(gb_ctx* c, const void* qkv, void* out, int B, int L, int D, int causal, cudaStream_t st);

int gb_launch_attn_fwd(gb_ctx* c, const void* qkv, void* out, int B, int L, int D, int causal,
                       cudaStream_t st) {
  if (B <= 0) return GB_OK;
  if (gb_attn_tc_enabled() && L <= 128 && D % kDh == 0) return gb_launch_attn_fwd_tc(c, qkv, out, B, L, D, causal, st);
  if (L < 1 || L > 96 || D % kDh != 0)
    return gb_fail(c, GB_ERR_ARG, "attention: L=%d (1..96) D=%d (multiple of 64) unsupported", L, D);
  const int lp = (L + 15) / 16;
  switch (lp) {
    case 1: return launch_fwd<2>(c, qkv, out, B, L, D, causal, st);
    case 2: return launch_fwd<4>(c, qkv, out, B, L, D, causal, st);
    case 3: return launch_fwd<6>(c, qkv, out, B, L, D, causal, st);
    case 4: return launch_fwd<8>(c, qkv, out, B, L, D, causal, st);
    case 5: return launch_fwd<10>(c, qkv, out, B, L, D, causal, st);
    default: return launch_fwd<12>(c, qkv, out, B, L, D, causal, st);
  }
}

int gb_launch_attn_bwd_tc(gb_ctx* c, const void* qkv, const void* dout, void* dqkv, int B, int L, int D, int causal,
                          cudaStream_t st);

int gb_launch_attn_bwd(gb_ctx* c, const void* qkv, const void* dout, void* dqkv, int B, int L,
                       int D, int causal, cudaStream_t st) {
  if (B <= 0) return GB_OK;
  if (gb_attn_tc_enabled() && L <= 96 && D % kDh == 0) return gb_launch_attn_bwd_tc(c, qkv, dout, dqkv, B, L, D, causal, st);
  if (L < 1 || L > 96 || D % kDh != 0)
    return gb_fail(c, GB_ERR_ARG, "attention: L=%d (1..96) D=%d (multiple of 64) unsupported", L, D);
  const int lp = (L + 15) / 16;
  switch (lp) {
    case 1: return launch_bwd<2>(c, qkv, dout, dqkv, B, L, D, causal, st);
    case 2: return launch_bwd<4>(c, qkv, dout, dqkv, B, L, D, causal, st);
    case 3: return launch_bwd<6>(c, qkv, dout, dqkv, B, L, D, causal, st);
    case 4: return launch_bwd<8>(c, qkv, dout, dqkv, B, L, D, causal, st);
    case 5: return launch_bwd<10>(c, qkv, dout, dqkv, B, L, D, causal, st);
    default: return launch_bwd<12>(c, qkv, dout, dqkv, B, L, D, causal, st);
  }
}

extern "C" int gb_attention_fwd(gb_ctx* c, const void* qkv, void* out, int B, int L, int D,
                                int causal, void* stream) {
  if (!c) return GB_ERR_ARG;
  gb_dev_guard dev_guard(c);
  if (!qkv || !out) return gb_fail(c, GB_ERR_ARG, "attention: null pointer");
  return gb_launch_attn_fwd(c, qkv, out, B, L, D, causal, (cudaStream_t)stream);
}

extern "C" int gb_attention_bwd(gb_ctx* c, const void* qkv, const void* dout, void* dqkv, int B,
                                int L, int D, int causal, void* stream) {
  if (!c) return GB_ERR_ARG;
  gb_dev_guard dev_guard(c);
  if (!qkv || !dout || !dqkv) return gb_fail(c, GB_ERR_ARG, "attention: null pointer");
  return gb_launch_attn_bwd(c, qkv, dout, dqkv, B, L, D, causal, (cudaStream_t)stream);
}
