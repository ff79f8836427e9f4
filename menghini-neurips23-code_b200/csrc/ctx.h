// Host-side context shared by all C-ABI entry points of libgripb200.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/gripb200.h"

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct gb_tower;  // tower.cu

// Key of a 2-D fp16 tensor map: everything cuTensorMapEncodeTiled is given.  A tensor map is pure metadata (address,
// extents, strides, box, swizzle), so a map encoded once for a key is valid for as long as the key is — whatever has
// happened to the memory in between.
struct gb_tmap_key {
  const void* ptr; uint64_t rows, cols, ld; uint32_t box_rows, box_cols;
  bool operator==(const gb_tmap_key& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows && box_cols == o.box_cols;
  }
};
struct gb_tmap_key_hash {
  size_t operator()(const gb_tmap_key& k) const {
    uint64_t h = reinterpret_cast<uint64_t>(k.ptr) * 0x9E3779B97F4A7C15ull;
    h ^= (k.rows + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    h ^= (k.cols * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2));
    h ^= (k.ld * 0x165667B19E3779F9ull + (h << 6) + (h >> 2));
    h ^= ((static_cast<uint64_t>(k.box_rows) << 32 | k.box_cols) + (h << 6) + (h >> 2));
    return static_cast<size_t>(h);
  }
};

struct gb_ctx {
  int device = 0;
  int num_sms = 148;
  std::string err;
  PFN_encodeTiled encode_tiled = nullptr;
  // cuTensorMapEncodeTiled costs ≈1 µs on the host and a GEMM launch needs three maps: at the reference's BATCH_SIZE
  // the towers are launch-bound and the same few hundred (buffer, shape) pairs come back every step
  std::unordered_map<gb_tmap_key, CUtensorMap, gb_tmap_key_hash> tmaps;
  uint64_t launches = 0;  // kernels launched through this ctx (bench.py reports it)
  // workspace owned by the ctx (grown on demand, never inside a timed region after warm-up)
  // One scratch arena per independent call family, so that the image tower, the text tower and the
  // pool scan may be in flight on different streams at the same time.
  enum { kWsVit = 0, kWsText = 1, kWsScan = 2, kWsTrain = 3, kWsCount = 4 };
  void* ws[kWsCount] = {nullptr, nullptr, nullptr, nullptr};
  size_t ws_bytes[kWsCount] = {0, 0, 0, 0};
  uint64_t ws_gen = 0;  // bumped whenever a workspace is re-allocated (captured CUDA graphs go stale)
  int sm_limit = 0;  // > 0: persistent GEMM grids use at most this many SMs (rounded down to pairs)
  gb_tower* vit = nullptr;
  gb_tower* text = nullptr;
  // optional per-launch timing of the two roofline kernels (bench.py): kind 0 = tcgen05 GEMM (work =
  // FLOPs), kind 1 = sim/softmax/argmax (work = algorithmic HBM bytes)
  bool prof_on = false;
  struct ProfRec { cudaEvent_t e0, e1; int kind; double work; int m, n, k; };
  std::vector<ProfRec> prof;
};

// RAII helper: records an event pair around one launch when profiling is on.
struct gb_prof_scope {
  gb_ctx* c; cudaStream_t st; bool on;
  gb_prof_scope(gb_ctx* c_, cudaStream_t st_, int kind, double work, int m = 0, int n = 0, int k = 0)
      : c(c_), st(st_), on(c_->prof_on) {
    if (!on) return;
    gb_ctx::ProfRec r;
    r.kind = kind; r.work = work; r.m = m; r.n = n; r.k = k;
    cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
    cudaEventRecord(r.e0, st);
    c->prof.push_back(r);
  }
  ~gb_prof_scope() { if (on) cudaEventRecord(c->prof.back().e1, st); }
};

// Every C-ABI entry point runs on its ctx's device whatever the caller's current device is, and leaves the
// caller's current device as it found it (one ctx per (process, device); several may be alive).
struct gb_dev_guard {
  int prev = -1, dev;
  explicit gb_dev_guard(const gb_ctx* c) : dev(c->device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev);
  }
  ~gb_dev_guard() {
    if (prev >= 0 && prev != dev) cudaSetDevice(prev);
  }
};

inline int gb_fail(gb_ctx* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  return code;
}

#define GB_CUDA(c, expr)                                                                  \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess)                                                                \
      return gb_fail((c), GB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                     __FILE__, __LINE__);                                                 \
  } while (0)

#define GB_LAUNCH_CHECK(c)                                                               \
  do {                                                                                   \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess)                                                               \
      return gb_fail((c), GB_ERR_CUDA, "kernel launch failed: %s (%s:%d)",               \
                     cudaGetErrorString(_e), __FILE__, __LINE__);                        \
    (c)->launches++;                                                                     \
  } while (0)

// 2-D fp16 row-major tensor map: box = {box_cols elements, box_rows}; 128 B swizzle for 64-column
// boxes, 64 B swizzle for 32-column boxes.
int gb_make_tmap_2d_f16(gb_ctx* c, CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols,
                        uint64_t ld_elems, uint32_t box_rows, uint32_t box_cols = 64);

// Epilogue warps of the CTA-pair GEMM (8 or 16; see gemm_tcgen05.cuh) and, with it, the columns per partial of
// gb_gemm_ln::stats_out — the columns one epilogue warp owns.
#ifndef GB_EPI_WARPS
#define GB_EPI_WARPS 8
#endif
#define GB_STAT_SEG (256 / (GB_EPI_WARPS / 4))

// LayerNorm-folding extras of a GEMM launch (see GemmParams): statistics consumed / produced.
struct gb_gemm_ln {
  const float* ln_stats = nullptr;  // [M][2] (μ·rstd, rstd) of A's rows → fold LayerNorm into this GEMM
  const float* ln_parts = nullptr;  // … or [nparts][M] float4 partials as a GEMM's stats_out left them
  int nparts = 0;                   //   (merged in the epilogue; K / nparts columns per segment)
  const float* col_sum = nullptr;   // [N]
  float* stats_out = nullptr;      // [N/GB_STAT_SEG][M] float4 shifted partial statistics of the output rows
};

// internal launchers shared between op-level and tower-level entry points
int gb_launch_gemm(gb_ctx* c, const void* A, int lda, const void* W, int ldw, const float* bias,
                   const void* resid, int ldr, void* out, int ldo, int M, int N, int K, int act,
                   int out_f32, cudaStream_t st, void* aux = nullptr, const struct gb_gemm_ln* ln = nullptr);
int gb_ws_reserve(gb_ctx* c, int slot, size_t bytes);
inline int gb_gemm_sms(const gb_ctx* c) { return (c->sm_limit > 0 && c->sm_limit < c->num_sms) ? c->sm_limit : c->num_sms; }
