/* libgripb200 — C ABI of the B200-native CLIP prompt-tuning / pseudolabel hot path.
 *
 * Drop-in boundary for BatsResearch/menghini-neurips23-code.  The reference has
 * no native layer: its hot path is Python calling torch ops through the third-party `clip`
 * package.  Each entry point below names the reference call it replaces (paths relative to the
 * reference root).  The Python classes in menghini-neurips23-code_b200/dropin/ bind these with
 * ctypes and keep the reference's class names / signatures.
 *
 * Conventions
 *   - plain pointers and sizes only; device pointers are raw CUDA addresses (tensor.data_ptr()).
 *   - every call is stream-ordered and asynchronous on `stream` (a cudaStream_t passed as void*).
 *   - return 0 on success, a negative gb_status otherwise; gb_last_error() gives the message.
 *   - never throws, never falls back to a CPU path: without a CUDA device gb_create() fails.
 *   - a ctx is bound to one device and is not thread-safe; distinct ctxs are independent.
 */
#ifndef GRIPB200_H_
#define GRIPB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gb_ctx gb_ctx;

typedef enum gb_status {
  GB_OK = 0,
  GB_ERR_CUDA = -1,     /* CUDA runtime / driver error (message has details) */
  GB_ERR_ARG = -2,      /* invalid argument (shape, alignment, null pointer) */
  GB_ERR_STATE = -3,    /* weights not loaded / workspace too small */
  GB_ERR_NO_DEVICE = -4 /* no sm_100 device: there is no CPU fallback */
} gb_status;

/* ---- context ----------------------------------------------------------------------------- */
int gb_create(gb_ctx** out, int device);
int gb_destroy(gb_ctx* ctx);
const char* gb_last_error(gb_ctx* ctx);
/* number of kernels launched through this ctx so far (bench.py's gpu_launches) */
uint64_t gb_launch_count(gb_ctx* ctx);
const char* gb_version(void);

/* ---- op level (unit-testable building blocks) -------------------------------------------- */

/* out[M,N] = epi(A[M,K] · W[N,K]^T): fp16 operands, fp32 accumulate (tcgen05 + TMA + TMEM).
 * Replaces torch nn.Linear / F.linear inside clip.model.ResidualAttentionBlock (in_proj, out_proj,
 * c_fc, c_proj), conv1 as an im2col GEMM (models/clip_encoders.py:131) and `x @ proj`
 * (models/clip_encoders.py:86-89,191-192).
 *   bias  : fp32 [N] or NULL.   act: 0 none, 1 QuickGELU (x*sigmoid(1.702x)).
 *   resid : fp16 [M,ldr] added after the activation, or NULL; may alias out (x += f(x)).
 *   out   : fp16 [M,ldo], or fp32 when out_f32 != 0.
 * Requirements: K % 64 == 0, N % 128 == 0, lda/ldw % 8 == 0, 16-byte aligned pointers. */
int gb_gemm_f16(gb_ctx* ctx, const void* A, int lda, const void* W, int ldw, const float* bias,
                const void* resid, int ldr, void* out, int ldo, int M, int N, int K, int act,
                int out_f32, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GRIPB200_H_ */
