// Row-wise kernels around the GEMMs: LayerNorm (CLIP's fp32-internal LayerNorm), im2col of the
// 32×32/stride-32 patch conv, token assembly (+ln_pre) for both towers, L2 row normalisation.
// All are single-pass, HBM-streaming kernels: one warp per row, 16-byte vector accesses.
#include "ctx.h"
#include "common.cuh"

using namespace gb;

namespace {

// ---------------------------------------------------------------------------------------------
// LayerNorm over D ∈ {512, 768} (D % 256 == 0).  y = (x-mean)/sqrt(var+eps)*g + b, fp32 math,
// two-pass variance like torch.native_layer_norm.  in_row = row_idx ? row_idx[r] : r*in_row_mul.
// Reference: clip.model.LayerNorm (x.float() → layer_norm → .type(orig)) used by ln_pre, ln_1,
// ln_2, ln_post (models/clip_encoders.py:157,189) and ln_final (:85).
// ---------------------------------------------------------------------------------------------
template <int D, typename OutT>
__global__ void __launch_bounds__(256)
layernorm_kernel(const __half* __restrict__ x, int ldx, const int32_t* __restrict__ row_idx,
                 int in_row_mul, const float* __restrict__ gamma, const float* __restrict__ beta,
                 OutT* __restrict__ y, int ldy, int rows, float eps) {
  constexpr int V = D / 256;  // 16-byte chunks per lane
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const size_t in_row = row_idx ? (size_t)row_idx[warp] : (size_t)warp * in_row_mul;
  const uint4* xr = reinterpret_cast<const uint4*>(x + in_row * ldx);
  float f[V * 8];
  float s = 0.f;
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const uint4 u = xr[v * 32 + lane];
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 a = __half22float2(h[t]);
      f[v * 8 + 2 * t] = a.x;
      f[v * 8 + 2 * t + 1] = a.y;
      s += a.x + a.y;
    }
  }
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < V * 8; ++i) {
    const float d = f[i] - mean;
    q += d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const int col = (v * 32 + lane) * 8;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + col));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + col + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + col));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + col + 4));
    float o[8];
    o[0] = (f[v * 8 + 0] - mean) * rstd * g0.x + b0.x;
    o[1] = (f[v * 8 + 1] - mean) * rstd * g0.y + b0.y;
    o[2] = (f[v * 8 + 2] - mean) * rstd * g0.z + b0.z;
    o[3] = (f[v * 8 + 3] - mean) * rstd * g0.w + b0.w;
    o[4] = (f[v * 8 + 4] - mean) * rstd * g1.x + b1.x;
    o[5] = (f[v * 8 + 5] - mean) * rstd * g1.y + b1.y;
    o[6] = (f[v * 8 + 6] - mean) * rstd * g1.z + b1.z;
    o[7] = (f[v * 8 + 7] - mean) * rstd * g1.w + b1.w;
    if constexpr (sizeof(OutT) == 2) {
      uint4 u;
      __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
      for (int t = 0; t < 4; ++t) h[t] = __floats2half2_rn(o[2 * t], o[2 * t + 1]);
      reinterpret_cast<uint4*>(y + (size_t)warp * ldy)[v * 32 + lane] = u;
    } else {
      float4* yo = reinterpret_cast<float4*>(y + (size_t)warp * ldy + col);
      yo[0] = make_float4(o[0], o[1], o[2], o[3]);
      yo[1] = make_float4(o[4], o[5], o[6], o[7]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// im2col for conv1 (kernel = stride = 32, no bias): out[(b*49 + py*7 + px), c*1024 + ky*32 + kx]
//   = img[b, c, py*32 + ky, px*32 + kx]   (fp32 or fp16 NCHW in → fp16 out).
// Because stride == kernel the re-index is a permutation: every input element is read once.
// Reference: self.conv1(x) + reshape/permute, models/clip_encoders.py:131-133.
// ---------------------------------------------------------------------------------------------
template <typename InT>
__global__ void __launch_bounds__(256)
im2col_patch32_kernel(const InT* __restrict__ img, __half* __restrict__ out, int B) {
  pdl_launch_dependents();
  // one thread per 8 consecutive kx: total = B*3*224*28 groups
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)B * 3 * 224 * 28;
  if (gid >= total) return;
  const int x8 = gid % 28;
  const int y = (gid / 28) % 224;
  const int c = (gid / (28 * 224)) % 3;
  const int b = gid / (28 * 224 * 3);
  const InT* src = img + (((size_t)b * 3 + c) * 224 + y) * 224 + x8 * 8;
  float v[8];
  if constexpr (sizeof(InT) == 4) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src));
    const float4 d = __ldg(reinterpret_cast<const float4*>(src) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = d.x; v[5] = d.y; v[6] = d.z; v[7] = d.w;
  } else {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(src));
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 a = __half22float2(h[t]);
      v[2 * t] = a.x; v[2 * t + 1] = a.y;
    }
  }
  const int py = y >> 5, ky = y & 31, px = (x8 * 8) >> 5, kx = (x8 * 8) & 31;
  uint4 o;
  __half2* h = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int t = 0; t < 4; ++t) h[t] = __floats2half2_rn(v[2 * t], v[2 * t + 1]);
  __half* dst = out + ((size_t)b * 49 + py * 7 + px) * 3072 + c * 1024 + ky * 32 + kx;
  *reinterpret_cast<uint4*>(dst) = o;
}

// Same re-index for raw uint8 pixels [B,3,224,224], with the reference's preprocessing tail fused in:
// torchvision ToTensor (x/255 in fp32) followed by Normalize with CLIP's mean / std ((x − mean)/std,
// clip.load's `_transform`; used by the reference at data/dataset.py:64-79), rounded to the fp16 GEMM operand.
// A pixel has 256 possible values per channel, so "the same bits as torch's three fp32 operations followed by the
// fp16 rounding" is a finite statement: fp16(fma(x, a_c, b_c)) with a_c = fp32(1/(255·std_c)), b_c = fp32(−mean_c/std_c)
// (constants formed in double) gives the identical fp16 value for all 3 × 256 inputs — checked exhaustively
// (tests/test_tokenizer.py::test_u8_normalise_fma_is_exhaustively_exact on the host, test_uint8_pixels_match_… on
// the device) — so the kernel needs neither the two IEEE divisions per pixel (ALU-bound at a third of the HBM rate)
// nor the per-block shared-memory table it used before (building 768 entries cost more than the 4 KB of pixels the
// block then looked up, and the random 2-byte lookups conflicted in the banks: 43 % of HBM peak in ncu).
// byte → float without the quarter-rate I2F: 0x4B000000 | byte is the float 2^23 + byte.
__global__ void __launch_bounds__(256)
im2col_patch32_u8_kernel(const uint8_t* __restrict__ img, __half* __restrict__ out, int B) {
  pdl_launch_dependents();
  // one thread per 16 consecutive kx: total = B*3*224*14 groups
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)B * 3 * 224 * 14;
  if (gid >= total) return;
  const int x16 = gid % 14;
  const int y = (gid / 14) % 224;
  const int c = (gid / (14 * 224)) % 3;
  const int b = gid / (14 * 224 * 3);
  constexpr float kA0 = (float)(1.0 / (255.0 * (double)0.26862954f)), kB0 = (float)(-(double)0.48145466f / (double)0.26862954f);
  constexpr float kA1 = (float)(1.0 / (255.0 * (double)0.26130258f)), kB1 = (float)(-(double)0.4578275f / (double)0.26130258f);
  constexpr float kA2 = (float)(1.0 / (255.0 * (double)0.27577711f)), kB2 = (float)(-(double)0.40821073f / (double)0.27577711f);
  const float ka = c == 0 ? kA0 : (c == 1 ? kA1 : kA2);
  const float kb = c == 0 ? kB0 : (c == 1 ? kB1 : kB2);
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(img + (((size_t)b * 3 + c) * 224 + y) * 224 + x16 * 16));
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
  uint4 o[2];
  __half2* h = reinterpret_cast<__half2*>(o);
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const uint32_t word = w[t >> 1];
    const float f0 = __uint_as_float(__byte_perm(word, 0x4B000000u, (t & 1) ? 0x7652 : 0x7650)) - 8388608.0f;
    const float f1 = __uint_as_float(__byte_perm(word, 0x4B000000u, (t & 1) ? 0x7653 : 0x7651)) - 8388608.0f;
    h[t] = __floats2half2_rn(__fmaf_rn(f0, ka, kb), __fmaf_rn(f1, ka, kb));
  }
  const int py = y >> 5, ky = y & 31, px = (x16 * 16) >> 5, kx = (x16 * 16) & 31;
  uint4* dst = reinterpret_cast<uint4*>(out + ((size_t)b * 49 + py * 7 + px) * 3072 + c * 1024 + ky * 32 + kx);
  dst[0] = o[0];
  dst[1] = o[1];
}

// ---------------------------------------------------------------------------------------------
// Vision token assembly fused with ln_pre (D = 768):
//   row l of image b:  l == 0      → class_embedding + pos[0]
//                      1 ≤ l ≤ P   → prefix[l-1]            (no positional embedding!)
//                      l > P       → patch[b, l-1-P] + pos[l-P]
//   x[b*L + l] = ln_pre(row)            (fp16 residual stream)
// Reference: models/clip_encoders.py:135-157 (prefix inserted after pos-emb, before ln_pre).
// ---------------------------------------------------------------------------------------------
// A warp owns token position l of kAsmRows consecutive images: the positional row (or the prefix row) and the
// ln_pre affine of its 24 columns per lane are read ONCE and stay in registers, so per row only the 1.5 KB patch row
// comes in and the 1.5 KB token row goes out (one warp per row re-read 9 KB of parameters per row through L2: 17 % of
// the HBM peak in ncu).  Per-row arithmetic is unchanged.
constexpr int kAsmRows = 8;
__global__ void __launch_bounds__(256)
vit_assemble_lnpre_kernel(const __half* __restrict__ patch, const float* __restrict__ cls,
                          const float* __restrict__ pos, const float* __restrict__ prefix, int P,
                          const float* __restrict__ gamma, const float* __restrict__ beta,
                          __half* __restrict__ x, int B, float eps, float* __restrict__ stats) {
  pdl_launch_dependents();
  constexpr int D = 768, V = 3;
  const int L = 50 + P;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int chunks = (B + kAsmRows - 1) / kAsmRows;
  if (wid >= chunks * L) return;
  const int l = wid % L, b0 = (wid / L) * kAsmRows;
  const int b1 = min(B, b0 + kAsmRows);
  // what every image shares at this position: the additive row and the affine
  float base[V * 8], gm[V * 8], bt[V * 8];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const int col = (v * 32 + lane) * 8;
    const float* src = (l >= 1 && l <= P) ? prefix + (size_t)(l - 1) * D : pos + (size_t)((l == 0) ? 0 : l - P) * D;
    const float4 p0 = __ldg(reinterpret_cast<const float4*>(src + col));
    const float4 p1 = __ldg(reinterpret_cast<const float4*>(src + col + 4));
    float* a = base + v * 8;
    a[0] = p0.x; a[1] = p0.y; a[2] = p0.z; a[3] = p0.w; a[4] = p1.x; a[5] = p1.y; a[6] = p1.z; a[7] = p1.w;
    if (l == 0) {
      const float4 c0 = __ldg(reinterpret_cast<const float4*>(cls + col));
      const float4 c1 = __ldg(reinterpret_cast<const float4*>(cls + col + 4));
      a[0] += c0.x; a[1] += c0.y; a[2] += c0.z; a[3] += c0.w; a[4] += c1.x; a[5] += c1.y; a[6] += c1.z; a[7] += c1.w;
    }
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + col)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + col + 4));
    const float4 e0 = __ldg(reinterpret_cast<const float4*>(beta + col)), e1 = __ldg(reinterpret_cast<const float4*>(beta + col + 4));
    float* gg = gm + v * 8;
    float* bb = bt + v * 8;
    gg[0] = g0.x; gg[1] = g0.y; gg[2] = g0.z; gg[3] = g0.w; gg[4] = g1.x; gg[5] = g1.y; gg[6] = g1.z; gg[7] = g1.w;
    bb[0] = e0.x; bb[1] = e0.y; bb[2] = e0.z; bb[3] = e0.w; bb[4] = e1.x; bb[5] = e1.y; bb[6] = e1.z; bb[7] = e1.w;
  }
  const bool has_patch = l > P;
  uint4 nxt[V];   // the next image's patch row is requested before this one is worked on
  if (has_patch) {
#pragma unroll
    for (int v = 0; v < V; ++v)
      nxt[v] = *reinterpret_cast<const uint4*>(patch + ((size_t)b0 * 49 + (l - 1 - P)) * D + (v * 32 + lane) * 8);
  }
  for (int b = b0; b < b1; ++b) {
  const int warp = b * L + l;   // token row
  float f[V * 8];
  uint4 cur[V];
#pragma unroll
  for (int v = 0; v < V; ++v) cur[v] = nxt[v];
  if (has_patch && b + 1 < b1) {
#pragma unroll
    for (int v = 0; v < V; ++v)
      nxt[v] = *reinterpret_cast<const uint4*>(patch + ((size_t)(b + 1) * 49 + (l - 1 - P)) * D + (v * 32 + lane) * 8);
  }
#pragma unroll
  for (int v = 0; v < V; ++v) {
#pragma unroll
    for (int i = 0; i < 8; ++i) f[v * 8 + i] = base[v * 8 + i];
    if (has_patch) {
      const uint4 u = cur[v];
      const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 e = __half22float2(h[t]);
        f[v * 8 + 2 * t] += e.x; f[v * 8 + 2 * t + 1] += e.y;
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < V * 8; ++i) s += f[i];
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < V * 8; ++i) { const float d = f[i] - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
  float st_sum = 0.f, st_sq = 0.f;
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const int col = (v * 32 + lane) * 8;
    uint4 u;
    __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float o0 = (f[v * 8 + 2 * t] - mean) * rstd * gm[v * 8 + 2 * t] + bt[v * 8 + 2 * t];
      const float o1 = (f[v * 8 + 2 * t + 1] - mean) * rstd * gm[v * 8 + 2 * t + 1] + bt[v * 8 + 2 * t + 1];
      h[t] = __floats2half2_rn(o0, o1);
      const float2 r = __half22float2(h[t]);
      f[v * 8 + 2 * t] = r.x;  // keep the stored (rounded) values for the statistics below
      f[v * 8 + 2 * t + 1] = r.y;
      st_sum += r.x + r.y;
    }
    *reinterpret_cast<uint4*>(x + (size_t)warp * D + col) = u;
  }
  // (μ·rstd, rstd) of the stored row, two-pass: the statistics of the first block's folded ln_1
  if (stats != nullptr) {
    const float m2 = warp_sum(st_sum) * (1.0f / D);
#pragma unroll
    for (int i = 0; i < V * 8; ++i) { const float d = f[i] - m2; st_sq = fmaf(d, d, st_sq); }
    const float r2 = rsqrtf(warp_sum(st_sq) * (1.0f / D) + eps);
    if (lane == 0) *reinterpret_cast<float2*>(stats + (size_t)warp * 2) = make_float2(m2 * r2, r2);
  }
  }  // images of this warp
}

// ---------------------------------------------------------------------------------------------
// Text token assembly (D = 512, ctx = 77):
//   x[c*77 + l] = (1 ≤ l ≤ P ? prefix[l-1] : token_embedding[ids[c,l]]) + pos[l]
// Reference: models/clip_encoders.py:63-74 (rows 1..P of the embedded prompt are overwritten by the
// learnable prefix, then the positional embedding is added to every row).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
text_assemble_kernel(const int32_t* __restrict__ ids, int ld_ids, const __half* __restrict__ tok_emb,
                     const float* __restrict__ pos, const float* __restrict__ prefix, int P,
                     __half* __restrict__ x, int C, int ctx_len, float* __restrict__ stats) {
  pdl_launch_dependents();
  constexpr int D = 512, V = 2;
  float st_sum = 0.f, st_sq = 0.f;
  float kept[V * 8];  // the stored (rounded) row, for the two-pass statistics
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= C * ctx_len) return;
  const int l = warp % ctx_len;
  const bool is_prefix = (l >= 1 && l <= P);
  const int id = ids[(size_t)(warp / ctx_len) * ld_ids + l];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const int col = (v * 32 + lane) * 8;
    float a[8];
    if (is_prefix) {
      const float4 p0 = __ldg(reinterpret_cast<const float4*>(prefix + (size_t)(l - 1) * D + col));
      const float4 p1 = __ldg(reinterpret_cast<const float4*>(prefix + (size_t)(l - 1) * D + col + 4));
      a[0] = p0.x; a[1] = p0.y; a[2] = p0.z; a[3] = p0.w; a[4] = p1.x; a[5] = p1.y; a[6] = p1.z; a[7] = p1.w;
    } else {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(tok_emb + (size_t)id * D + col));
      const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 e = __half22float2(h[t]);
        a[2 * t] = e.x; a[2 * t + 1] = e.y;
      }
    }
    const float4 q0 = __ldg(reinterpret_cast<const float4*>(pos + (size_t)l * D + col));
    const float4 q1 = __ldg(reinterpret_cast<const float4*>(pos + (size_t)l * D + col + 4));
    a[0] += q0.x; a[1] += q0.y; a[2] += q0.z; a[3] += q0.w; a[4] += q1.x; a[5] += q1.y; a[6] += q1.z; a[7] += q1.w;
    uint4 o;
    __half2* h = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      h[t] = __floats2half2_rn(a[2 * t], a[2 * t + 1]);
      const float2 r = __half22float2(h[t]);
      kept[v * 8 + 2 * t] = r.x;
      kept[v * 8 + 2 * t + 1] = r.y;
      st_sum += r.x + r.y;
    }
    *reinterpret_cast<uint4*>(x + (size_t)warp * D + col) = o;
  }
  if (stats != nullptr) {  // (μ·rstd, rstd) of the stored row, two-pass
    const float m2 = warp_sum(st_sum) * (1.0f / D);
#pragma unroll
    for (int i = 0; i < V * 8; ++i) { const float d = kept[i] - m2; st_sq = fmaf(d, d, st_sq); }
    const float r2 = rsqrtf(warp_sum(st_sq) * (1.0f / D) + 1e-5f);
    if (lane == 0) *reinterpret_cast<float2*>(stats + (size_t)warp * 2) = make_float2(m2 * r2, r2);
  }
}

// ---------------------------------------------------------------------------------------------
// Row L2 normalisation of [rows, 512] fp32 features → fp16 (and/or fp32) unit rows.
// Reference: `x / x.norm(dim=-1, keepdim=True)` at every logits site, e.g.
// methods/semi_supervised_learning/textual_prompt.py:98-103 and CLIP.forward.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
l2norm512_kernel(const float* __restrict__ x, __half* __restrict__ y16, float* __restrict__ y32,
                 int rows) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)warp * 512);
  float4 v[4];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[i] = xr[i * 32 + lane];
    s += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
  }
  const float inv = 1.0f / sqrtf(warp_sum(s));
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 o = make_float4(v[i].x * inv, v[i].y * inv, v[i].z * inv, v[i].w * inv);
    if (y32) reinterpret_cast<float4*>(y32 + (size_t)warp * 512)[i * 32 + lane] = o;
    if (y16) {
      uint2 u;
      __half2* h = reinterpret_cast<__half2*>(&u);
      h[0] = __floats2half2_rn(o.x, o.y);
      h[1] = __floats2half2_rn(o.z, o.w);
      reinterpret_cast<uint2*>(y16 + (size_t)warp * 512)[i * 32 + lane] = u;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm backward w.r.t. the input only (γ, β are frozen):
//   x̂ = (x−μ)·rstd,  g = dy∘γ,  dx = rstd·(g − mean(g) − x̂·mean(g∘x̂))
// dy row r pairs with x row xr = row_idx ? row_idx[r] : r*in_row_mul; the result is written (or, with
// accumulate, added: the residual branch) to dx row xr.  fp16 streams, fp32 math.
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const __half* __restrict__ dy, int lddy, const __half* __restrict__ x, int ldx,
                     const int32_t* __restrict__ row_idx, int in_row_mul,
                     const float* __restrict__ gamma, __half* __restrict__ dx, int lddx, int rows,
                     int accumulate, float eps) {
  pdl_launch_dependents();
  constexpr int V = D / 256;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const size_t xr = row_idx ? (size_t)row_idx[warp] : (size_t)warp * in_row_mul;
  const uint4* xp = reinterpret_cast<const uint4*>(x + xr * ldx);
  const uint4* dp = reinterpret_cast<const uint4*>(dy + (size_t)warp * lddy);
  float f[V * 8], g[V * 8];
  float s = 0.f;
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const uint4 u = xp[v * 32 + lane];
    const uint4 w = dp[v * 32 + lane];
    const __half2* h = reinterpret_cast<const __half2*>(&u);
    const __half2* hd = reinterpret_cast<const __half2*>(&w);
    const int col = (v * 32 + lane) * 8;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + col));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + col + 4));
    const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 a = __half22float2(h[t]);
      const float2 d = __half22float2(hd[t]);
      f[v * 8 + 2 * t] = a.x;
      f[v * 8 + 2 * t + 1] = a.y;
      g[v * 8 + 2 * t] = d.x * gm[2 * t];
      g[v * 8 + 2 * t + 1] = d.y * gm[2 * t + 1];
      s += a.x + a.y;
    }
  }
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < V * 8; ++i) { const float d = f[i] - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
  float m1 = 0.f, m2 = 0.f;
#pragma unroll
  for (int i = 0; i < V * 8; ++i) {
    f[i] = (f[i] - mean) * rstd;  // x̂
    m1 += g[i];
    m2 += g[i] * f[i];
  }
  m1 = warp_sum(m1) * (1.0f / D);
  m2 = warp_sum(m2) * (1.0f / D);
  uint4* op = reinterpret_cast<uint4*>(dx + xr * lddx);
#pragma unroll
  for (int v = 0; v < V; ++v) {
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = rstd * (g[v * 8 + i] - m1 - f[v * 8 + i] * m2);
    if (accumulate) {
      const uint4 u = op[v * 32 + lane];
      const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 a = __half22float2(h[t]);
        o[2 * t] += a.x;
        o[2 * t + 1] += a.y;
      }
    }
    uint4 u;
    __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
    for (int t = 0; t < 4; ++t) h[t] = __floats2half2_rn(o[2 * t], o[2 * t + 1]);
    op[v * 32 + lane] = u;
  }
}

// ---------------------------------------------------------------------------------------------
// Gradient of the learnable prompt rows: dprefix[p,:] = inv_scale · Σ_s d(row 1+p of sample s).
// Vision (ln_pre != 0): the rows went through ln_pre with input prefix[p] (models/clip_encoders.py:
// 148-157), so each sample's row gradient is first pulled back through that LayerNorm.
// Text: the prefix rows are the raw embeddings (+pos), gradient = plain sum over the C prompts
// (models/clip_encoders.py:67).  One CTA per prompt row, warps stride over samples.
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256)
prefix_grad_kernel(const __half* __restrict__ dx, int L, int S, const float* __restrict__ prefix,
                   const float* __restrict__ gamma, int ln_pre, float inv_scale,
                   float* __restrict__ dprefix, float eps) {
  constexpr int V = D / 32;  // elements per lane, strided by 32
  __shared__ float red[8][D];
  const int p = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float xh[V], gm[V], acc[V];
  float rstd = 1.f;
#pragma unroll
  for (int i = 0; i < V; ++i) { acc[i] = 0.f; xh[i] = 0.f; gm[i] = 1.f; }
  if (ln_pre) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) { xh[i] = prefix[(size_t)p * D + i * 32 + lane]; s += xh[i]; }
    const float mean = warp_sum(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) { xh[i] -= mean; q += xh[i] * xh[i]; }
    rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
#pragma unroll
    for (int i = 0; i < V; ++i) { xh[i] *= rstd; gm[i] = gamma[i * 32 + lane]; }
  }
  for (int s = warp; s < S; s += 8) {
    const __half* r = dx + ((size_t)s * L + 1 + p) * D;
    float g[V];
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      g[i] = __half2float(r[i * 32 + lane]) * gm[i];
      m1 += g[i];
      m2 += g[i] * xh[i];
    }
    if (ln_pre) {
      m1 = warp_sum(m1) * (1.0f / D);
      m2 = warp_sum(m2) * (1.0f / D);
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] += rstd * (g[i] - m1 - xh[i] * m2);
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] += g[i];
    }
  }
#pragma unroll
  for (int i = 0; i < V; ++i) red[warp][i * 32 + lane] = acc[i];
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += 256) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][c];
    dprefix[(size_t)p * D + c] = t * inv_scale;
  }
}

// out16[r, :] = fp16(scale · in32[r, :])   (loss-scaled entry of the fp16 gradient stream)
__global__ void __launch_bounds__(256)
scale_f32_to_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, size_t n,
                        float scale) {
  const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(in + i);
    uint2 u;
    __half2* h = reinterpret_cast<__half2*>(&u);
    h[0] = __floats2half2_rn(v.x * scale, v.y * scale);
    h[1] = __floats2half2_rn(v.z * scale, v.w * scale);
    *reinterpret_cast<uint2*>(out + i) = u;
  } else {
    for (size_t j = i; j < n; ++j) out[j] = __float2half_rn(in[j] * scale);
  }
}

inline int warps_grid(long long rows) { return (int)((rows * 32 + 255) / 256); }

}  // namespace

// ---- internal launchers (used by tower.cu) ------------------------------------------------------
int gb_launch_layernorm(gb_ctx* c, const void* x, int ldx, const int32_t* row_idx, int in_row_mul,
                        const float* gamma, const float* beta, void* y, int ldy, int rows, int D,
                        int out_f32, cudaStream_t st) {
  if (rows <= 0) return GB_OK;
  if (D != 512 && D != 768) return gb_fail(c, GB_ERR_ARG, "layernorm: D must be 512 or 768");
  const int grid = warps_grid(rows);
  const __half* xi = reinterpret_cast<const __half*>(x);
  if (D == 768) {
    if (out_f32) layernorm_kernel<768, float><<<grid, 256, 0, st>>>(xi, ldx, row_idx, in_row_mul, gamma, beta, (float*)y, ldy, rows, 1e-5f);
    else layernorm_kernel<768, __half><<<grid, 256, 0, st>>>(xi, ldx, row_idx, in_row_mul, gamma, beta, (__half*)y, ldy, rows, 1e-5f);
  } else {
    if (out_f32) layernorm_kernel<512, float><<<grid, 256, 0, st>>>(xi, ldx, row_idx, in_row_mul, gamma, beta, (float*)y, ldy, rows, 1e-5f);
    else layernorm_kernel<512, __half><<<grid, 256, 0, st>>>(xi, ldx, row_idx, in_row_mul, gamma, beta, (__half*)y, ldy, rows, 1e-5f);
  }
  GB_LAUNCH_CHECK(c);
  return GB_OK;
}

int gb_launch_im2col(gb_ctx* c, const void* img, int img_fmt, void* out, int B, cudaStream_t st) {
  if (B <= 0) return GB_OK;
  if (img_fmt < 0 || img_fmt > GB_IMG_U8) return gb_fail(c, GB_ERR_ARG, "im2col: unknown image format %d", img_fmt);
  if (reinterpret_cast<uintptr_t>(img) & 15) return gb_fail(c, GB_ERR_ARG, "im2col: image pointer must be 16-byte aligned");
  const size_t total = (size_t)B * 3 * 224 * (img_fmt == GB_IMG_U8 ? 14 : 28);
  const int grid = (int)((total + 255) / 256);
  if (img_fmt == GB_IMG_U8) im2col_patch32_u8_kernel<<<grid, 256, 0, st>>>((const uint8_t*)img, (__half*)out, B);
  else if (img_fmt == GB_IMG_F32) im2col_patch32_kernel<float><<<grid, 256, 0, st>>>((const float*)img, (__half*)out, B);
  else im2col_patch32_kernel<__half><<<grid, 256, 0, st>>>((const __half*)img, (__half*)out, B);
  GB_LAUNCH_CHECK(c);
  return GB_OK;
}

int gb_launch_vit_assemble(gb_ctx* c, const void* patch, const float* cls, const float* pos,
                           const float* prefix, int P, const float* gamma, const float* beta,
                           void* x, int B, cudaStream_t st, float* stats) {
  if (B <= 0) return GB_OK;
  vit_assemble_lnpre_kernel<<<warps_grid((long long)((B + kAsmRows - 1) / kAsmRows) * (50 + P)), 256, 0, st>>>(
      (const __half*)patch, cls, pos, prefix, P, gamma, beta, (__half*)x, B, 1e-5f, stats);
  GB_LAUNCH_CHECK(c);
  return GB_OK;
}

__global__ void eot_rows_kernel(const int32_t* __restrict__ eot, int32_t* __restrict__ rows, int C,
                                int Lt) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) rows[c] = c * Lt + min(eot[c], Lt - 1);
}

int gb_launch_eot_rows(gb_ctx* c, const int32_t* eot, int32_t* rows, int C, int Lt, cudaStream_t st) {
  if (C <= 0) return GB_OK;
  eot_rows_kernel<<<(C + 127) / 128, 128, 0, st>>>(eot, rows, C, Lt);
  GB_LAUNCH_CHECK(c);
  return GB_OK;
}

int gb_launch_text_assemble(gb_ctx* c, const int32_t* ids, int ld_ids, const void* tok_emb,
                            const float* pos, const float* prefix, int P, void* x, int C, int ctx_len,
                            cudaStream_t st, float* stats) {
  if (C <= 0) return GB_OK;
  text_assemble_kernel<<<warps_grid((long long)C * ctx_len), 256, 0, st>>>(
      ids, ld_ids, (const __half*)tok_emb, pos, prefix, P, (__half*)x, C, ctx_len, stats);
  GB_LAUNCH_CHECK(c);
  return GB_OK;
}

int gb_launch_l2norm512(gb_ctx* c, const float* x, void* y16, float* y32, int rows,
                        cudaStream_t st) {
  if (rows <= 0) return GB_OK;
  l2norm512_kernel<<<warps_grid(rows), 256, 0, st>>>(x, (__half*)y16, y32, rows);
  GB_LAUNCH_CHECK(c);
  return GB_OK;
}

int gb_launch_layernorm_bwd(gb_ctx* c, const void* dy, int lddy, const void* x, int ldx,
                            const int32_t* row_idx, int in_row_mul, const float* gamma, void* dx,
                            int lddx, int rows, int D, int accumulate, cudaStream_t st) {
  if (rows <= 0) return GB_OK;
  if (D != 512 && D != 768) return gb_fail(c, GB_ERR_ARG, "layernorm_bwd: D must be 512 or 768");
  const int grid = warps_grid(rows);
  if (D == 768)
    layernorm_bwd_kernel<768><<<grid, 256, 0, st>>>((const __half*)dy, lddy, (const __half*)x, ldx, row_idx, in_row_mul, gamma, (__half*)dx, lddx, rows, accumulate, 1e-5f);
  else
    layernorm_bwd_kernel<512><<<grid, 256, 0, st>>>((const __half*)dy, lddy, (const __half*)x, ldx, row_idx, in_row_mul, gamma, (__half*)dx, lddx, rows, accumulate, 1e-5f);
  GB_LAUNCH_CHECK(c);
  return GB_OK;
}

int gb_launch_prefix_grad(gb_ctx* c, const void* dx, int L, int S, int P, int D, const float* prefix,
                          const float* gamma, int ln_pre, float inv_scale, float* dprefix,
                          cudaStream_t st) {
  if (P <= 0) return GB_OK;
  if (D == 768)
    prefix_grad_kernel<768><<<P, 256, 0, st>>>((const __half*)dx, L, S, prefix, gamma, ln_pre, inv_scale, dprefix, 1e-5f);
  else if (D == 512)
    prefix_grad_kernel<512><<<P, 256, 0, st>>>((const __half*)dx, L, S, prefix, gamma, ln_pre, inv_scale, dprefix, 1e-5f);
  else
    return gb_fail(c, GB_ERR_ARG, "prefix_grad: D must be 512 or 768");
  GB_LAUNCH_CHECK(c);
  return GB_OK;
}

int gb_launch_scale_f32_to_f16(gb_ctx* c, const float* in, void* out, size_t n, float scale,
                               cudaStream_t st) {
  if (n == 0) return GB_OK;
  const size_t threads = (n + 3) / 4;
  scale_f32_to_f16_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(in, (__half*)out, n, scale);
  GB_LAUNCH_CHECK(c);
  return GB_OK;
}

// ---- C ABI ----------------------------------------------------------------------------------------
extern "C" int gb_layernorm_f16(gb_ctx* c, const void* x, int ldx, const int32_t* row_idx,
                                int in_row_mul, const float* gamma, const float* beta, void* y,
                                int ldy, int rows, int D, int out_f32, void* stream) {
  if (!c) return GB_ERR_ARG;
  gb_dev_guard dev_guard(c);
  if (!x || !gamma || !beta || !y) return gb_fail(c, GB_ERR_ARG, "layernorm: null pointer");
  return gb_launch_layernorm(c, x, ldx, row_idx, in_row_mul, gamma, beta, y, ldy, rows, D, out_f32,
                             (cudaStream_t)stream);
}

extern "C" int gb_l2norm512(gb_ctx* c, const float* x, void* y16, float* y32, int rows,
                            void* stream) {
  if (!c) return GB_ERR_ARG;
  gb_dev_guard dev_guard(c);
  if (!x || (!y16 && !y32)) return gb_fail(c, GB_ERR_ARG, "l2norm: null pointer");
  return gb_launch_l2norm512(c, x, y16, y32, rows, (cudaStream_t)stream);
}
