// Prompt-tuning step glue on the device: the cosine-logit cross-entropy of the reference's
// training loops and the SGD update of the prompt parameters, without host round trips.
//
// Reference: methods/semi_supervised_learning/textual_prompt.py:93-135 — L2-normalise text and image
// features, logits = logit_scale.exp()·I·Tᵀ, nn.CrossEntropyLoss, backward, optimizer.step(); the FPL
// variants weight two groups of samples differently (balance_param, textual_fpl.py:123-165), which is a
// per-sample coefficient here.  The learning-rate rule is utils/schedulers.py:36-65 (WarmupCosineSchedule).
#include <math.h>

#include "common.cuh"
#include "ctx.h"

using namespace gb;

namespace {

__device__ __forceinline__ float to_f32(__half x) { return __half2float(x); }
__device__ __forceinline__ float to_f32(float x) { return x; }

// One warp per image row: logits against the C normalised prompts, log-softmax, this row's loss term and
// d loss / d logits.  tn = T / |T| is recomputed from T by every warp's lanes on the fly (C·512 floats,
// L2-resident); all sums run in a fixed order (deterministic).
//   loss_i = −coef_i · log softmax(z_i)[y_i],   dz_ij = coef_i · (softmax(z_i)_j − [j == y_i])
template <typename TI>
__global__ void __launch_bounds__(256)
ce_rows_kernel(const TI* __restrict__ imfn, const float* __restrict__ tn, const int32_t* __restrict__ labels,
               const float* __restrict__ coef, float coef_all, float scale, int B, int C,
               float* __restrict__ dlogits, float* __restrict__ loss_rows, int32_t* __restrict__ pred) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  // this row's features: lane holds 16 of the 512 values (two 16-byte pieces)
  float f[16];
  if constexpr (sizeof(TI) == 2) {
    const uint4* src = reinterpret_cast<const uint4*>(imfn + (size_t)row * 512);
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const uint4 u = src[p * 32 + lane];
      const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 a = __half22float2(h[t]);
        f[p * 8 + 2 * t] = a.x;
        f[p * 8 + 2 * t + 1] = a.y;
      }
    }
  } else {  // fp32 unit rows, same element ↔ lane mapping
    const float4* src = reinterpret_cast<const float4*>(imfn + (size_t)row * 512);
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const float4 a = src[(p * 32 + lane) * 2], b = src[(p * 32 + lane) * 2 + 1];
      f[p * 8 + 0] = a.x; f[p * 8 + 1] = a.y; f[p * 8 + 2] = a.z; f[p * 8 + 3] = a.w;
      f[p * 8 + 4] = b.x; f[p * 8 + 5] = b.y; f[p * 8 + 6] = b.z; f[p * 8 + 7] = b.w;
    }
  }
  float* z = dlogits + (size_t)row * C;
  float mx = -INFINITY;
  int am = 0;
  for (int j = 0; j < C; ++j) {
    const float4* t4 = reinterpret_cast<const float4*>(tn + (size_t)j * 512);
    float s = 0.f;
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const float4 a = __ldg(t4 + (p * 32 + lane) * 2);
      const float4 b = __ldg(t4 + (p * 32 + lane) * 2 + 1);
      s += f[p * 8 + 0] * a.x + f[p * 8 + 1] * a.y + f[p * 8 + 2] * a.z + f[p * 8 + 3] * a.w +
           f[p * 8 + 4] * b.x + f[p * 8 + 5] * b.y + f[p * 8 + 6] * b.z + f[p * 8 + 7] * b.w;
    }
    s = warp_sum(s) * scale;
    if (lane == 0) z[j] = s;
    if (s > mx) { mx = s; am = j; }
  }
  __syncwarp();
  float sum = 0.f;
  for (int j = lane; j < C; j += 32) sum += expf(z[j] - mx);
  sum = warp_sum(sum);
  const float lse = mx + logf(sum);
  const int y = labels[row];
  // a label outside [0, C) (torch raises for it) poisons the loss and the gradient instead of leaving
  // loss_rows[row] unwritten
  const bool y_ok = y >= 0 && y < C;
  const float cf = y_ok ? (coef ? coef[row] : coef_all) : __int_as_float(0x7fc00000);
  if (!y_ok && lane == 0) loss_rows[row] = cf;
  for (int j = lane; j < C; j += 32) {
    const float zj = z[j];
    if (j == y) loss_rows[row] = cf * (lse - zj);
    z[j] = cf * (expf(zj - lse) - (j == y ? 1.f : 0.f));
  }
  if (pred && lane == 0) pred[row] = am;
}

// tn = T / |T| (fp32 [C,512]) and 1/|T| per row: one warp per prompt.
__global__ void __launch_bounds__(256)
text_unit_kernel(const float* __restrict__ T, float* __restrict__ tn, float* __restrict__ inv_norm, int C) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= C) return;
  const float4* src = reinterpret_cast<const float4*>(T + (size_t)row * 512);
  float4 v[4];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[i] = src[i * 32 + lane];
    s += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
  }
  const float inv = 1.0f / sqrtf(warp_sum(s));
#pragma unroll
  for (int i = 0; i < 4; ++i)
    reinterpret_cast<float4*>(tn + (size_t)row * 512)[i * 32 + lane] =
        make_float4(v[i].x * inv, v[i].y * inv, v[i].z * inv, v[i].w * inv);
  if (lane == 0) inv_norm[row] = inv;
}

// dTn[j,d] = scale·Σ_i dz[i,j]·I[i,d] in two deterministic stages: block (j, s) adds up the rows of segment s
// (kCeSeg consecutive rows, index order) into part[s][j][d]; ce_text_grad_kernel then adds the segments in index
// order.  (One block per prompt walking all B rows was 256 µs at B = 1966, C = 10: ten CTAs, a serial chain of
// dependent loads each.)
constexpr int kCeSeg = 64;
template <typename TI>
__global__ void __launch_bounds__(512)
ce_text_part_kernel(const float* __restrict__ dlogits, const TI* __restrict__ imfn, int B, int C,
                    float* __restrict__ part) {
  const int j = blockIdx.x, s = blockIdx.y, d = threadIdx.x;
  const int i0 = s * kCeSeg, i1 = min(B, i0 + kCeSeg);
  float acc = 0.f;
#pragma unroll 8
  for (int i = i0; i < i1; ++i) acc = fmaf(__ldg(dlogits + (size_t)i * C + j), to_f32(imfn[(size_t)i * 512 + d]), acc);
  part[((size_t)s * C + j) * 512 + d] = acc;
}

// One block (512 threads) per prompt j: dTn_j from the segment partials, then back through the normalisation:
// dT_j = (dTn_j − tn_j·⟨tn_j, dTn_j⟩)/|T_j|.  Block 0 also adds up the per-row losses in index order.
__global__ void __launch_bounds__(512)
ce_text_grad_kernel(const float* __restrict__ part, int nseg, const float* __restrict__ tn,
                    const float* __restrict__ inv_norm, const float* __restrict__ loss_rows, float scale, int B,
                    int C, float* __restrict__ dT, float* __restrict__ loss) {
  __shared__ float red[16];
  const int j = blockIdx.x, d = threadIdx.x;
  float acc = 0.f;
  for (int s = 0; s < nseg; ++s) acc += part[((size_t)s * C + j) * 512 + d];
  acc *= scale;
  const float t = tn[(size_t)j * 512 + d];
  float dot = warp_sum(t * acc);
  if ((d & 31) == 0) red[d >> 5] = dot;
  __syncthreads();
  dot = 0.f;
#pragma unroll
  for (int w = 0; w < 16; ++w) dot += red[w];
  dT[(size_t)j * 512 + d] = (acc - t * dot) * inv_norm[j];
  if (j == 0 && loss) {
    __syncthreads();
    float s = 0.f;
    for (int i = d; i < B; i += 512) s += loss_rows[i];
    s = warp_sum(s);
    if ((d & 31) == 0) red[d >> 5] = s;
    __syncthreads();
    if (d == 0) {
      float tot = 0.f;
      for (int w = 0; w < 16; ++w) tot += red[w];
      *loss = tot;
    }
  }
}

// in = I / |I| (fp32 [B,512]) and 1/|I| per row: one warp per image (visual_prompt.py:123-124).
__global__ void __launch_bounds__(256)
image_unit_kernel(const float* __restrict__ I, float* __restrict__ in, float* __restrict__ inv_norm, int B) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const float4* src = reinterpret_cast<const float4*>(I + (size_t)row * 512);
  float4 v[4];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[i] = src[i * 32 + lane];
    s += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
  }
  const float inv = 1.0f / sqrtf(warp_sum(s));
#pragma unroll
  for (int i = 0; i < 4; ++i)
    reinterpret_cast<float4*>(in + (size_t)row * 512)[i * 32 + lane] =
        make_float4(v[i].x * inv, v[i].y * inv, v[i].z * inv, v[i].w * inv);
  if (lane == 0) inv_norm[row] = inv;
}

// One warp per image i: dIn_i = scale·Σ_j dz[i,j]·tn_j (classes in index order), then back through the
// normalisation: dI_i = (dIn_i − in_i·⟨in_i, dIn_i⟩)/|I_i|.  Lane l holds columns 4·(l + 32·q) … +3, q < 4.
__global__ void __launch_bounds__(256)
ce_image_grad_kernel(const float* __restrict__ dlogits, const float* __restrict__ in, const float* __restrict__ tn,
                     const float* __restrict__ inv_norm, float scale, int B, int C, float* __restrict__ dI) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  float4 acc[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* dz = dlogits + (size_t)row * C;
  for (int j = 0; j < C; ++j) {
    const float w = dz[j];
    const float4* t4 = reinterpret_cast<const float4*>(tn + (size_t)j * 512);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 t = __ldg(t4 + q * 32 + lane);
      acc[q].x = fmaf(w, t.x, acc[q].x); acc[q].y = fmaf(w, t.y, acc[q].y);
      acc[q].z = fmaf(w, t.z, acc[q].z); acc[q].w = fmaf(w, t.w, acc[q].w);
    }
  }
  const float4* i4 = reinterpret_cast<const float4*>(in + (size_t)row * 512);
  float4 u[4];
  float dot = 0.f;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    u[q] = i4[q * 32 + lane];
    acc[q].x *= scale; acc[q].y *= scale; acc[q].z *= scale; acc[q].w *= scale;
    dot += u[q].x * acc[q].x + u[q].y * acc[q].y + u[q].z * acc[q].z + u[q].w * acc[q].w;
  }
  dot = warp_sum(dot);
  const float inv = inv_norm[row];
#pragma unroll
  for (int q = 0; q < 4; ++q)
    reinterpret_cast<float4*>(dI + (size_t)row * 512)[q * 32 + lane] =
        make_float4((acc[q].x - u[q].x * dot) * inv, (acc[q].y - u[q].y * dot) * inv,
                    (acc[q].z - u[q].z * dot) * inv, (acc[q].w - u[q].w * dot) * inv);
}

// Σ_i loss_rows[i] in the order ce_text_grad_kernel's block 0 uses (same bits with or without a text gradient).
__global__ void __launch_bounds__(512) loss_sum_kernel(const float* __restrict__ loss_rows, int B, float* __restrict__ loss) {
  __shared__ float red[16];
  const int d = threadIdx.x;
  float s = 0.f;
  for (int i = d; i < B; i += 512) s += loss_rows[i];
  s = warp_sum(s);
  if ((d & 31) == 0) red[d >> 5] = s;
  __syncthreads();
  if (d == 0) {
    float tot = 0.f;
    for (int w = 0; w < 16; ++w) tot += red[w];
    *loss = tot;
  }
}

// torch.optim.SGD (dampening 0, no Nesterov) with the learning rate of utils/schedulers.py:36-65 evaluated
// on the device from the epoch counter the caller keeps there:  g ← g + wd·p;  b ← μ·b + g (b ← g on the
// first step);  p ← p − lr·(μ ? b : g).
__global__ void __launch_bounds__(256)
sgd_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ mom, size_t n, float lr,
                const float* __restrict__ lr_dev, float momentum, float weight_decay, int first) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (lr_dev) lr = *lr_dev;  // a captured CUDA graph keeps following the schedule
  float gi = fmaf(weight_decay, p[i], g[i]);
  if (momentum != 0.f) {
    const float b = first ? gi : fmaf(momentum, mom[i], gi);
    mom[i] = b;
    gi = b;
  }
  p[i] -= lr * gi;
}

}  // namespace

extern "C" int gb_ce_text_grad(gb_ctx* c, const void* imfn16, const float* text, const int32_t* labels,
                               const float* coef, float logit_scale_exp, int B, int C, float* dtext,
                               float* loss, int32_t* pred, void* stream) {
  if (!c) return GB_ERR_ARG;
  gb_dev_guard dev_guard(c);
  if (!imfn16 || !text || !labels || !dtext || B <= 0 || C <= 0)
    return gb_fail(c, GB_ERR_ARG, "ce_text_grad: bad arguments (B=%d C=%d)", B, C);
  cudaStream_t st = (cudaStream_t)stream;
  // scratch: tn [C,512] | inv_norm [C] | dlogits [B,C] | loss_rows [B] | segment partials [nseg][C][512]
  const int nseg = (B + kCeSeg - 1) / kCeSeg;
  const size_t need = ((size_t)C * 512 + C + (size_t)B * C + B + 16 + (size_t)nseg * C * 512) * 4 + 64;
  int rc = gb_ws_reserve(c, gb_ctx::kWsTrain, need);
  if (rc) return rc;
  float* tn = reinterpret_cast<float*>(c->ws[gb_ctx::kWsTrain]);
  float* inv_norm = tn + (size_t)C * 512;
  float* dz = inv_norm + ((C + 3) & ~3);
  float* loss_rows = dz + (((size_t)B * C + 3) & ~(size_t)3);
  text_unit_kernel<<<(C * 32 + 255) / 256, 256, 0, st>>>(text, tn, inv_norm, C);
  GB_LAUNCH_CHECK(c);
  ce_rows_kernel<__half><<<(int)(((size_t)B * 32 + 255) / 256), 256, 0, st>>>(
      (const __half*)imfn16, tn, labels, coef, 1.0f / (float)B, logit_scale_exp, B, C, dz, loss_rows, pred);
  GB_LAUNCH_CHECK(c);
  float* part = loss_rows + ((B + 3) & ~3);
  ce_text_part_kernel<__half><<<dim3(C, nseg), 512, 0, st>>>(dz, (const __half*)imfn16, B, C, part);
  GB_LAUNCH_CHECK(c);
  ce_text_grad_kernel<<<C, 512, 0, st>>>(part, nseg, tn, inv_norm, loss_rows, logit_scale_exp, B, C, dtext, loss);
  GB_LAUNCH_CHECK(c);
  return GB_OK;
}

extern "C" int gb_ce_image_grad(gb_ctx* c, const float* image, const float* text, const int32_t* labels,
                                const float* coef, float logit_scale_exp, int B, int C, float* dimage,
                                float* dtext, float* loss, int32_t* pred, void* stream) {
  if (!c) return GB_ERR_ARG;
  gb_dev_guard dev_guard(c);
  if (!image || !text || !labels || (!dimage && !dtext) || B <= 0 || C <= 0)
    return gb_fail(c, GB_ERR_ARG, "ce_image_grad: bad arguments (B=%d C=%d)", B, C);
  cudaStream_t st = (cudaStream_t)stream;
  // scratch: tn [C,512] | inv_t [C] | in [B,512] | inv_i [B] | dlogits [B,C] | loss_rows [B] | segment partials
  const int nseg = (B + kCeSeg - 1) / kCeSeg;
  const size_t need = ((size_t)C * 512 + C + 4 + (size_t)B * 512 + B + 4 + (size_t)B * C + 4 + B + 16 +
                       (dtext ? (size_t)nseg * C * 512 : 0)) * 4 + 64;
  int rc = gb_ws_reserve(c, gb_ctx::kWsTrain, need);
  if (rc) return rc;
  float* tn = reinterpret_cast<float*>(c->ws[gb_ctx::kWsTrain]);
  float* inv_t = tn + (size_t)C * 512;
  float* in = inv_t + ((C + 3) & ~3);
  float* inv_i = in + (size_t)B * 512;
  float* dz = inv_i + ((B + 3) & ~3);
  float* loss_rows = dz + (((size_t)B * C + 3) & ~(size_t)3);
  text_unit_kernel<<<(C * 32 + 255) / 256, 256, 0, st>>>(text, tn, inv_t, C);
  GB_LAUNCH_CHECK(c);
  image_unit_kernel<<<(int)(((size_t)B * 32 + 255) / 256), 256, 0, st>>>(image, in, inv_i, B);
  GB_LAUNCH_CHECK(c);
  ce_rows_kernel<float><<<(int)(((size_t)B * 32 + 255) / 256), 256, 0, st>>>(
      in, tn, labels, coef, 1.0f / (float)B, logit_scale_exp, B, C, dz, loss_rows, pred);
  GB_LAUNCH_CHECK(c);
  if (dimage) {
    ce_image_grad_kernel<<<(int)(((size_t)B * 32 + 255) / 256), 256, 0, st>>>(dz, in, tn, inv_i, logit_scale_exp, B,
                                                                             C, dimage);
    GB_LAUNCH_CHECK(c);
  }
  // the text-side kernel also adds up the per-row losses; without a text gradient a 1-prompt launch of it
  // would not do: run it on all prompts into the scratch tail only when asked, else sum the loss alone
  if (dtext) {
    float* part = loss_rows + ((B + 3) & ~3);
    ce_text_part_kernel<float><<<dim3(C, nseg), 512, 0, st>>>(dz, in, B, C, part);
    GB_LAUNCH_CHECK(c);
    ce_text_grad_kernel<<<C, 512, 0, st>>>(part, nseg, tn, inv_t, loss_rows, logit_scale_exp, B, C, dtext, loss);
    GB_LAUNCH_CHECK(c);
  } else if (loss) {
    loss_sum_kernel<<<1, 512, 0, st>>>(loss_rows, B, loss);
    GB_LAUNCH_CHECK(c);
  }
  return GB_OK;
}

extern "C" double gb_warmup_cosine_lr(double base_lr, int warmup_steps, int t_total, int step) {
  // utils/schedulers.py:54-65 (cycles = 0.5): factor = step / max(1, warmup) during warm-up, then
  // max(0, 0.5·(1 + cos(π·progress))) with progress = (step − warmup) / max(1, t_total − warmup); in double,
  // like the Python it restates
  if (step < warmup_steps) return base_lr * ((double)step / (double)(warmup_steps > 1 ? warmup_steps : 1));
  const int rest = t_total - warmup_steps;
  const double progress = (double)(step - warmup_steps) / (double)(rest > 1 ? rest : 1);
  const double f = 0.5 * (1.0 + cos(M_PI * 0.5 * 2.0 * progress));
  return base_lr * (f > 0.0 ? f : 0.0);
}

extern "C" int gb_sgd_step(gb_ctx* c, float* param, const float* grad, float* momentum_buf, long long n, float lr,
                           const float* lr_dev, float momentum, float weight_decay, int first_step, void* stream) {
  if (!c) return GB_ERR_ARG;
  gb_dev_guard dev_guard(c);
  if (!param || !grad || n <= 0 || (momentum != 0.f && !momentum_buf))
    return gb_fail(c, GB_ERR_ARG, "sgd_step: bad arguments");
  sgd_step_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(param, grad, momentum_buf, (size_t)n,
                                                                                 lr, lr_dev, momentum, weight_decay,
                                                                                 first_step);
  GB_LAUNCH_CHECK(c);
  return GB_OK;
}
