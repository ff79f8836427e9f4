// Context management + TMA descriptor encoding for libgripb200.
#include "ctx.h"

extern "C" const char* gb_version(void) { return "gripb200 0.1 (sm_100a)"; }

extern "C" int gb_create(gb_ctx** out, int device) {
  if (!out) return GB_ERR_ARG;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) {
    cudaGetLastError();
    return GB_ERR_NO_DEVICE;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return GB_ERR_CUDA;
  if (prop.major != 10) return GB_ERR_NO_DEVICE;  // tcgen05 kernels are sm_100a-only
  gb_ctx* c = new gb_ctx();
  c->device = device;
  gb_dev_guard dev_guard(c);  // the caller's current device is restored on return
  cudaFree(0);
  c->num_sms = prop.multiProcessorCount;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) !=
          cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || fn == nullptr) {
    delete c;
    return GB_ERR_CUDA;
  }
  c->encode_tiled = reinterpret_cast<PFN_encodeTiled>(fn);
  *out = c;
  return GB_OK;
}

void gb_tower_free(gb_tower* t);

extern "C" int gb_destroy(gb_ctx* c) {
  if (!c) return GB_ERR_ARG;
  gb_dev_guard dev_guard(c);
  cudaSetDevice(c->device);
  for (int i = 0; i < gb_ctx::kWsCount; ++i)
    if (c->ws[i]) cudaFree(c->ws[i]);
  gb_tower_free(c->vit);
  gb_tower_free(c->text);
  delete c;
  return GB_OK;
}

extern "C" const char* gb_last_error(gb_ctx* c) { return c ? c->err.c_str() : "null ctx"; }
extern "C" uint64_t gb_launch_count(gb_ctx* c) { return c ? c->launches : 0; }

int gb_ws_reserve(gb_ctx* c, int slot, size_t bytes) {
  if (bytes <= c->ws_bytes[slot]) return GB_OK;
  GB_CUDA(c, cudaDeviceSynchronize());
  if (c->ws[slot]) GB_CUDA(c, cudaFree(c->ws[slot]));
  c->ws[slot] = nullptr;
  c->ws_bytes[slot] = 0;
  GB_CUDA(c, cudaMalloc(&c->ws[slot], bytes));
  c->ws_bytes[slot] = bytes;
  c->ws_gen++;
  return GB_OK;
}

extern "C" uint64_t gb_workspace_generation(gb_ctx* c) { return c ? c->ws_gen : 0; }

extern "C" int gb_set_sm_limit(gb_ctx* c, int sms) {
  if (!c || sms < 0) return GB_ERR_ARG;
  c->sm_limit = sms;
  return GB_OK;
}

int gb_make_tmap_2d_f16(gb_ctx* c, CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols,
                        uint64_t ld_elems, uint32_t box_rows, uint32_t box_cols) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  if (box_cols != 64 && box_cols != 32) return gb_fail(c, GB_ERR_ARG, "tensor map: box of %u columns", box_cols);
  const gb_tmap_key key{ptr, rows, cols, ld_elems, box_rows, box_cols};
  auto hit = c->tmaps.find(key);
  if (hit != c->tmaps.end()) {
    *out = hit->second;
    return GB_OK;
  }
  CUresult r = c->encode_tiled(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims,
                               strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return gb_fail(c, GB_ERR_CUDA,
                   "cuTensorMapEncodeTiled failed (%d) ptr=%p rows=%llu cols=%llu ld=%llu box=%u",
                   (int)r, ptr, (unsigned long long)rows, (unsigned long long)cols,
                   (unsigned long long)ld_elems, box_rows);
  if (c->tmaps.size() >= 4096) c->tmaps.clear();   // callers that stream through ever-new buffers: bounded memory
  c->tmaps.emplace(key, *out);
  return GB_OK;
}

extern "C" int gb_profile_begin(gb_ctx* c) {
  if (!c) return GB_ERR_ARG;
  gb_dev_guard dev_guard(c);
  for (auto& r : c->prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  c->prof.clear();
  c->prof_on = true;
  return GB_OK;
}

extern "C" int gb_profile_launches(gb_ctx* c, gb_profile_launch* out, int cap) {
  if (!c || (cap > 0 && !out)) return GB_ERR_ARG;
  gb_dev_guard dev_guard(c);
  if (cudaDeviceSynchronize() != cudaSuccess) return GB_ERR_CUDA;
  int n = 0;
  for (auto& r : c->prof) {
    if (n < cap) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, r.e0, r.e1);
      out[n].kind = r.kind; out[n].m = r.m; out[n].n = r.n; out[n].k = r.k;
      out[n].ms = ms; out[n].work = r.work;
    }
    ++n;
  }
  return n;
}

extern "C" int gb_profile_end(gb_ctx* c, gb_profile_stats* out, int kinds) {
  if (!c || !out || kinds <= 0) return GB_ERR_ARG;
  gb_dev_guard dev_guard(c);
  c->prof_on = false;
  GB_CUDA(c, cudaDeviceSynchronize());
  for (int k = 0; k < kinds; ++k) { out[k].launches = 0; out[k].ms = 0; out[k].work = 0; }
  for (auto& r : c->prof) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    if (r.kind < kinds) { out[r.kind].launches++; out[r.kind].ms += ms; out[r.kind].work += r.work; }
    cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
  }
  c->prof.clear();
  return GB_OK;
}
