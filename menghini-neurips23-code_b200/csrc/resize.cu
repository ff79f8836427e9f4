// Pillow's bicubic resize + CLIP's centre crop on the device, bit for bit (8-bit RGB).
//
// Reference: clip.load's `_transform` = Resize(224, BICUBIC) → CenterCrop(224) → …, applied on the host per image at
// data/dataset.py:64-79 and utils/clip_pseudolabels.py:56-57.  Pillow's resampler (src/libImaging/Resample.c) is integer
// arithmetic once the per-output-coordinate windows and 22-bit fixed-point weights are known; those come from the host
// (menghini-neurips23-code_b200/utils/pil_resample.py::coeffs, the same double operations in the same order, checked
// against Pillow by tests/test_pil_resample.py).  Two passes like ImagingResample: horizontal into a uint8 intermediate
// (only the 224 cropped columns of the rows the vertical pass will read), then vertical into planar [3,224,224] pixels —
// what `clip.preprocess_u8()` returns, ready for the image tower's uint8 patch gather.
// HBM-bound integer / byte work: H·W·3 bytes in, 224·224·3 out per image; one thread per output pixel (3 channels).
#include "common.cuh"
#include "ctx.h"

namespace {

constexpr int kPrec = 22;   // PRECISION_BITS = 32 − 8 − 2

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= kPrec;
  return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// tmp[img][r − row0][o][c] for output columns left … left+223 of rows row0 … row0+rows−1
__global__ void __launch_bounds__(224)
resize_h_kernel(const uint8_t* __restrict__ src, const int64_t* __restrict__ src_off, int H, int W,
                const int32_t* __restrict__ bx, const int32_t* __restrict__ kx, int ksx, int left, int row0, int rows,
                uint8_t* __restrict__ tmp) {
  const int o = threadIdx.x, r = blockIdx.x, img = blockIdx.y;
  const uint8_t* image = src_off ? src + src_off[img] : src + (size_t)img * H * W * 3;
  const uint8_t* line = image + (size_t)(row0 + r) * W * 3;
  uint8_t* dst = tmp + (((size_t)img * rows + r) * 224 + o) * 3;
  if (bx == nullptr) {   // the width is kept: crop only
    const uint8_t* p = line + (size_t)(left + o) * 3;
    dst[0] = p[0]; dst[1] = p[1]; dst[2] = p[2];
    return;
  }
  const int x0 = bx[2 * (left + o)], n = bx[2 * (left + o) + 1];
  const int32_t* k = kx + (size_t)(left + o) * ksx;
  int s0 = 1 << (kPrec - 1), s1 = s0, s2 = s0;
  const uint8_t* p = line + (size_t)x0 * 3;
  for (int x = 0; x < n; ++x) {
    const int w = __ldg(k + x);
    s0 += p[3 * x] * w; s1 += p[3 * x + 1] * w; s2 += p[3 * x + 2] * w;
  }
  dst[0] = clip8(s0); dst[1] = clip8(s1); dst[2] = clip8(s2);
}

// out[slot][c][y][x] from tmp rows; slot = out_index ? out_index[img] : img
__global__ void __launch_bounds__(224)
resize_v_kernel(const uint8_t* __restrict__ tmp, int rows, int row0, const int32_t* __restrict__ by,
                const int32_t* __restrict__ ky, int ksy, int top, const int32_t* __restrict__ out_index,
                uint8_t* __restrict__ out) {
  const int x = threadIdx.x, y = blockIdx.x, img = blockIdx.y;
  const size_t slot = out_index ? (size_t)out_index[img] : (size_t)img;
  uint8_t* dst = out + slot * 3 * 224 * 224 + (size_t)y * 224 + x;
  const uint8_t* base = tmp + (size_t)img * rows * 224 * 3 + (size_t)x * 3;
  if (by == nullptr) {   // the height is kept: crop only (row0 == top)
    const uint8_t* p = base + (size_t)y * 224 * 3;
    dst[0] = p[0]; dst[224 * 224] = p[1]; dst[2 * 224 * 224] = p[2];
    return;
  }
  const int y0 = by[2 * (top + y)] - row0, n = by[2 * (top + y) + 1];
  const int32_t* k = ky + (size_t)(top + y) * ksy;
  int s0 = 1 << (kPrec - 1), s1 = s0, s2 = s0;
  for (int t = 0; t < n; ++t) {
    const uint8_t* p = base + (size_t)(y0 + t) * 224 * 3;
    const int w = __ldg(k + t);
    s0 += p[0] * w; s1 += p[1] * w; s2 += p[2] * w;
  }
  dst[0] = clip8(s0); dst[224 * 224] = clip8(s1); dst[2 * 224 * 224] = clip8(s2);
}

// Two independent 64-bit multiply-sum checksums per row of `row_bytes` bytes (a multiple of 8): word i is multiplied by
// odd, index-dependent 64-bit constants and everything is added modulo 2^64 (order-free, hence deterministic).  What
// engine.frozen_image_features recognises an already encoded image by.
__global__ void __launch_bounds__(256) checksum128_kernel(const uint64_t* __restrict__ data, size_t words_per_row,
                                                          uint64_t* __restrict__ out) {
  __shared__ uint64_t red[2][8];
  const uint64_t* row = data + (size_t)blockIdx.x * words_per_row;
  uint64_t h0 = 0, h1 = 0;
  for (size_t i = threadIdx.x; i < words_per_row; i += 256) {
    const uint64_t x = row[i];
    const uint64_t m0 = ((i + 1) * 0x9E3779B97F4A7C15ull + 0x7F4A7C15ull) | 1ull;
    const uint64_t m1 = ((i + 1) * 0xC2B2AE3D27D4EB4Full + 0x165667B1ull) | 1ull;
    h0 += x * m0;
    h1 += (x ^ (x >> 29)) * m1;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    h0 += __shfl_xor_sync(0xffffffffu, h0, o);
    h1 += __shfl_xor_sync(0xffffffffu, h1, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = h0; red[1][threadIdx.x >> 5] = h1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t a = 0, b = 0;
    for (int w = 0; w < 8; ++w) { a += red[0][w]; b += red[1][w]; }
    out[2 * blockIdx.x] = a;
    out[2 * blockIdx.x + 1] = b;
  }
}

}  // namespace

// out[r] = two 64-bit checksums of row r (`row_bytes` bytes each, a multiple of 8; rows contiguous, 8-byte aligned).
extern "C" int gb_checksum128(gb_ctx* c, const void* data, int rows, size_t row_bytes, uint64_t* out, void* stream) {
  if (!c) return GB_ERR_ARG;
  gb_dev_guard dev_guard(c);
  if (rows <= 0) return GB_OK;
  if (!data || !out || row_bytes == 0 || (row_bytes & 7) || (reinterpret_cast<uintptr_t>(data) & 7))
    return gb_fail(c, GB_ERR_ARG, "checksum128: rows of a multiple of 8 bytes, 8-byte aligned");
  checksum128_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint64_t*>(data), row_bytes / 8, out);
  GB_LAUNCH_CHECK(c);
  return GB_OK;
}

extern "C" size_t gb_resize_tmp_bytes(int n, int rows) {
  return n <= 0 || rows <= 0 ? 0 : (size_t)n * rows * 224 * 3;
}

// n images of one size, uint8 [H,W,3] each (RGB, interleaved), image i at src + src_off[i] (src_off null: packed
// [n,H,W,3]) → out[slot] = uint8 [3,224,224].  bx/kx [nw,2] / [nw,ksx]:
// windows and weights of the horizontal pass (null: the width is already nw = W), `left` the crop origin in the resized
// width; by/ky/top likewise for the vertical pass; row0/rows: the source rows the vertical pass reads
// (by[top].first … by[top+223].first + count, or top … top+224 without a vertical pass); tmp: gb_resize_tmp_bytes(n, rows).
extern "C" int gb_resize_bicubic_crop_u8(gb_ctx* c, const uint8_t* src, const int64_t* src_off, int n, int H, int W,
                                         const int32_t* bx,
                                         const int32_t* kx, int ksx, int left, const int32_t* by, const int32_t* ky,
                                         int ksy, int top, int row0, int rows, const int32_t* out_index, uint8_t* tmp,
                                         uint8_t* out, void* stream) {
  if (!c) return GB_ERR_ARG;
  gb_dev_guard dev_guard(c);
  if (n <= 0) return GB_OK;
  if (!src || !tmp || !out || H <= 0 || W <= 0 || rows <= 0 || row0 < 0 || row0 + rows > H || left < 0 || top < 0 ||
      (bx && (!kx || ksx <= 0)) || (by && (!ky || ksy <= 0)) || (!bx && left + 224 > W) || (!by && (rows != 224 || row0 != top)))
    return gb_fail(c, GB_ERR_ARG, "resize: bad arguments (n=%d H=%d W=%d rows=%d row0=%d left=%d top=%d)", n, H, W, rows,
                   row0, left, top);
  cudaStream_t st = (cudaStream_t)stream;
  resize_h_kernel<<<dim3(rows, n), 224, 0, st>>>(src, src_off, H, W, bx, kx, ksx, left, row0, rows, tmp);
  GB_LAUNCH_CHECK(c);
  resize_v_kernel<<<dim3(224, n), 224, 0, st>>>(tmp, rows, row0, by, ky, ksy, top, out_index, out);
  GB_LAUNCH_CHECK(c);
  return GB_OK;
}
