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

#include <stddef.h>
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
/* Cap the persistent GEMM grids at `sms` SMs (0 = all).  Leaves a few SMs free for small kernels of a
 * second stream (e.g. the text tower running beside the image tower). */
int gb_set_sm_limit(gb_ctx* ctx, int sms);

/* Per-launch CUDA-event timing of the two roofline kernels, for bench.py: between _begin and _end
 * every tcgen05 GEMM launch (kind 0, work = 2*M*N*K FLOP) and every similarity/softmax/argmax launch
 * (kind 1, work = algorithmic HBM bytes) is bracketed by an event pair on its own stream. */
typedef struct gb_profile_stats { uint64_t launches; double ms; double work; } gb_profile_stats;
typedef struct gb_profile_launch { int kind, m, n, k; double ms; double work; } gb_profile_launch;
int gb_profile_begin(gb_ctx* ctx);
/* Per-launch records since _begin (call before _end); returns the number of launches recorded. */
int gb_profile_launches(gb_ctx* ctx, gb_profile_launch* out, int cap);
int gb_profile_end(gb_ctx* ctx, gb_profile_stats* out, int kinds);

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

/* y = LayerNorm(x) with fp32 statistics (clip.model.LayerNorm: ln_pre/ln_1/ln_2/ln_post/ln_final,
 * models/clip_encoders.py:85,157,189).  x fp16 [*,ldx]; input row r = row_idx ? row_idx[r] :
 * r*in_row_mul (gathers the CLS / EOT rows); D in {512,768}; y fp16 or fp32 [rows,ldy]. */
int gb_layernorm_f16(gb_ctx* ctx, const void* x, int ldx, const int32_t* row_idx, int in_row_mul,
                     const float* gamma, const float* beta, void* y, int ldy, int rows, int D,
                     int out_f32, void* stream);

/* Row L2 normalisation `x / x.norm(dim=-1, keepdim=True)` of fp32 [rows,512] features
 * (methods/semi_supervised_learning/textual_prompt.py:98-103) → fp16 and/or fp32 unit rows. */
int gb_l2norm512(gb_ctx* ctx, const float* x, void* y16, float* y32, int rows, void* stream);

/* Pillow's bicubic Image.resize + CLIP's centre crop on the device, bit for bit, for n 8-bit RGB images of one size:
 * image i = uint8 [H,W,3] at src + src_off[i] (src_off null: packed [n,H,W,3]) → out[slot] uint8 [3,224,224]
 * (slot = out_index ? out_index[i] : i), i.e. what the reference's
 * transform (clip.load's Resize(224, BICUBIC) → CenterCrop(224), data/dataset.py:64-79, utils/clip_pseudolabels.py:56)
 * yields before ToTensor / Normalize — which gb_vit_forward(…, GB_IMG_U8) applies.  bx int32 [nw,2] (first source
 * index, count) and kx int32 [nw,ksx] (22-bit fixed-point weights) describe the horizontal pass (null: width kept),
 * by / ky / ksy the vertical one; left / top are the crop origin in the resized image; row0 / rows the source rows the
 * vertical pass reads; tmp holds gb_resize_tmp_bytes(n, rows).  The tables come from
 * menghini-neurips23-code_b200/utils/pil_resample.py::coeffs (checked against Pillow, tests/test_pil_resample.py). */
size_t gb_resize_tmp_bytes(int n, int rows);
int gb_resize_bicubic_crop_u8(gb_ctx* ctx, const uint8_t* src, const int64_t* src_off, int n, int H, int W,
                              const int32_t* bx,
                              const int32_t* kx, int ksx, int left, const int32_t* by, const int32_t* ky, int ksy,
                              int top, int row0, int rows, const int32_t* out_index, uint8_t* tmp, uint8_t* out,
                              void* stream);

/* Two independent 64-bit multiply-sum checksums per row (out uint64 [rows][2]) of `rows` contiguous rows of `row_bytes`
 * bytes (a multiple of 8, 8-byte aligned): how the host side recognises an image the frozen tower has already encoded
 * (the reference re-encodes the same training images in every epoch, methods/semi_supervised_learning/
 * textual_prompt.py:99-103). */
int gb_checksum128(gb_ctx* ctx, const void* data, int rows, size_t row_bytes, uint64_t* out, void* stream);

/* softmax(Q K^T / 8 [+ causal mask]) V per (sample, head) on the packed in-proj output
 * qkv fp16 [B*L, 3D] → out fp16 [B*L, D]; head dim 64; forward L <= 128, backward L <= 96 (CLIP's sequences are
 * 50 + P <= 66 vision tokens and <= 77 text tokens).  nn.MultiheadAttention core of
 * clip.model.ResidualAttentionBlock.  _bwd: data gradient dqkv from dout (weights are frozen).  Both run on the
 * tcgen05 tensor cores (csrc/attn_tc.cu, csrc/attn_bwd_tc.cu; GB_ATTN_LEGACY=1 in the environment selects the
 * round-1 mma.sync kernels, L <= 96). */
int gb_attention_fwd(gb_ctx* ctx, const void* qkv, void* out, int B, int L, int D, int causal,
                     void* stream);
int gb_attention_bwd(gb_ctx* ctx, const void* qkv, const void* dout, void* dqkv, int B, int L,
                     int D, int causal, void* stream);

/* ---- towers ------------------------------------------------------------------------------- */

/* One ResidualAttentionBlock.  Weights fp16 in nn.Linear layout [out,in]; biases and LayerNorm
 * parameters fp32 (ln*_g / ln*_b are always the original γ, β).  *_t are [in,out] transposed copies used by the prompt-gradient pass
 * (may be NULL when only forward is needed). */
typedef struct gb_block_weights {
  const float* ln1_g; const float* ln1_b;
  const void* w_qkv;  const float* b_qkv;   /* [3D,D], [3D] */
  const void* w_o;    const float* b_o;     /* [D,D],  [D]  */
  const float* ln2_g; const float* ln2_b;
  const void* w_fc;   const float* b_fc;    /* [4D,D], [4D] */
  const void* w_proj; const float* b_proj;  /* [D,4D], [D]  */
  const void* w_qkv_t; const void* w_o_t; const void* w_fc_t; const void* w_proj_t;
  /* LayerNorm folding (optional, both or neither).  When s_qkv / s_fc are non-NULL, w_qkv / b_qkv hold
   * W∘γ1 and b + W·β1 (w_fc / b_fc: γ2, β2) and s_*[n] = Σ_k W'[n,k] in fp32: the forward then never
   * materialises ln_1(x) / ln_2(x); the GEMM epilogue applies the per-row mean and rstd instead.
   * The *_t copies stay the transposes of the ORIGINAL weights (the gradient pass uses γ itself). */
  const float* s_qkv; const float* s_fc;
} gb_block_weights;

/* Frozen CLIP ViT-B/32 image tower (clip.model.VisionTransformer as re-wired by
 * models/clip_encoders.py:105-194).  The table is copied; the device buffers stay caller-owned. */
typedef struct gb_vit_weights {
  int width, layers, heads, out_dim;        /* 768, 12, 12, 512 */
  const void* conv_w;                        /* fp16 [768, 3*32*32] (conv1.weight flattened) */
  const float* cls;                          /* [768] class_embedding */
  const float* pos;                          /* [50,768] positional_embedding */
  const float* ln_pre_g;  const float* ln_pre_b;
  const float* ln_post_g; const float* ln_post_b;
  const void* proj_t;                        /* fp16 [512,768] = proj^T */
  const void* proj;                          /* fp16 [768,512] = proj (gradient pass) or NULL */
  const gb_block_weights* blocks;            /* [layers] */
} gb_vit_weights;

/* Frozen CLIP text tower (token_embedding, positional_embedding, transformer, ln_final,
 * text_projection: models/clip_encoders.py:29-37). */
typedef struct gb_text_weights {
  int width, layers, heads, out_dim, ctx_len, vocab; /* 512, 12, 8, 512, 77, 49408 */
  const void* tok_emb;                       /* fp16 [vocab,512] */
  const float* pos;                          /* [77,512] */
  const float* ln_final_g; const float* ln_final_b;
  const void* proj_t;                        /* fp16 [512,512] = text_projection^T */
  const void* proj;                          /* fp16 [512,512] = text_projection or NULL */
  const gb_block_weights* blocks;
} gb_text_weights;

int gb_vit_set_weights(gb_ctx* ctx, const gb_vit_weights* w);
int gb_text_set_weights(gb_ctx* ctx, const gb_text_weights* w);

/* Bytes of the activation tape a forward must fill for a later prompt-gradient pass over
 * `samples` sequences of `L` tokens of width D (0 on bad arguments). */
size_t gb_tape_bytes(int samples, int L, int D, int layers);

/* CustomVisionTransformer.forward (models/clip_encoders.py:123-194) / encode_image when P == 0:
 * conv1 patch embed, CLS + positional embedding, P prefix rows inserted after CLS (no positional
 * embedding), ln_pre, 12 blocks, ln_post(CLS) @ proj.
 *   img    : [B,3,224,224] NCHW; img_f32 selects the element type: GB_IMG_F16 (0), GB_IMG_F32 (1) —
 *            already normalised, what the reference's DataLoader yields — or GB_IMG_U8 (2): raw 0..255
 *            pixels of the resized / centre-cropped image, ToTensor + Normalize(CLIP mean, std) applied
 *            on the device with torch's own fp32 operation order (bit-identical features, 4x fewer
 *            host→device bytes)
 *   prefix : fp32 [P,768] or NULL
 *   feat   : fp32 [B,512] un-normalised (what the reference modules return) or NULL
 *   featn  : fp16 [B,512] L2-normalised rows (input of gb_sim_softmax_argmax) or NULL
 *   tape   : gb_tape_bytes(B, 50+P, 768, 12) bytes or NULL (inference) */
enum { GB_IMG_F16 = 0, GB_IMG_F32 = 1, GB_IMG_U8 = 2 };
int gb_vit_forward(gb_ctx* ctx, const void* img, int img_f32, const float* prefix, int B, int P,
                   float* feat, void* featn, void* tape, void* stream);
/* dprefix fp32 [P,768] = d loss / d prefix given dfeat fp32 [B,512] (autograd through
 * CustomImageEncoder w.r.t. ImagePrefixModel.prefix, models/prompts_models.py:52-61). */
int gb_vit_backward_prefix(gb_ctx* ctx, const float* dfeat, const float* prefix, int B, int P,
                           const void* tape, float* dprefix, void* stream);

/* CustomTextEncoder.forward after tokenisation (models/clip_encoders.py:63-89) / encode_text when
 * P == 0: token embedding, rows 1..P overwritten by the prefix, + positional embedding, 12 causal
 * blocks, ln_final, EOT-row gather, @ text_projection.
 *   ids : int32 [C, ld_ids] token ids; eot : int32 [C] = ids.argmax(-1)
 *   Lt  : positions processed, max(eot)+1 <= Lt <= 77.  Causality makes positions after EOT
 *         irrelevant to the EOT row, so Lt < 77 is exact, not an approximation. */
int gb_text_forward(gb_ctx* ctx, const int32_t* ids, int ld_ids, const int32_t* eot,
                    const float* prefix, int C, int P, int Lt, float* feat, void* featn, void* tape,
                    void* stream);
int gb_text_backward_prefix(gb_ctx* ctx, const float* dfeat, const int32_t* eot, int C, int P,
                            int Lt, const void* tape, float* dprefix, void* stream);

/* ---- pool scan: similarity + softmax + argmax + leaderboard ------------------------------- */

/* logits = scale * F T^T, probs = softmax(logits), pred = argmax: CLIP.forward + softmax + argmax of
 * utils/clip_pseudolabels.py:59-65 for all N images in one HBM pass.
 *   F fp16 [N,512], T fp16 [C,512], both with unit rows; C <= 512.  Up to 128 classes the prototypes stay
 *   resident in shared memory and F is read once; wider class sets (SUN397, CUB-200) run in chunks of 128
 *   classes with one extra pass over F per chunk and phase (partial statistics, merge, probabilities).
 *   mode 0: pred = argmax(probs) (clip_pseudolabels.py:63); 1: argmax(logits) (textual_fpl.py:228)
 *   pred int32 [N]; p_pred fp32 [N] = probs[i,pred[i]]; probs fp32 [N,C] or NULL. */
int gb_sim_softmax_argmax(gb_ctx* ctx, const void* F, const void* T, float scale, int N, int C,
                          int mode, int32_t* pred, float* p_pred, float* probs, void* stream);

/* The sequential per-class leaderboard of utils/clip_pseudolabels.py:46-101 (and the 9
 * assign_pseudo_labels copies), exactly: arrival-order fill, strict `<` against the LAST entry,
 * sort-and-truncate on admission, spill of rejected images into every other board.
 * `state` is caller-owned device memory of gb_leaderboard_state_bytes(C,k) bytes; it is
 * relocatable, so a shard can hand it to the shard that owns the next index range.
 * rank[i] orders images like their path strings (ties in sorted()); NULL = index order.
 * _update feeds rows [row_begin,row_end) of probs/pred (local row r = global image idx0 + r). */
size_t gb_leaderboard_state_bytes(int C, int k);
int gb_leaderboard_init(gb_ctx* ctx, void* state, int C, int k, void* stream);
int gb_leaderboard_update(gb_ctx* ctx, void* state, int C, int k, const float* probs,
                          const int32_t* pred, const int32_t* rank, int row_begin, int row_end,
                          int idx0, int prefilter, void* stream);
/* out_idx int32 [C,k] (list order, -1 padded), out_len int32 [C], out_p fp32 [C,k] or NULL. */
int gb_leaderboard_export(gb_ctx* ctx, const void* state, int C, int k, int32_t* out_idx,
                          int32_t* out_len, float* out_p, void* stream);
/* Fused scan of rows [0,N): gb_sim_softmax_argmax with the leaderboard pre-filter in its epilogue,
 * exact replay of the surviving rows between chunks.  Replaces the loop
 * utils/clip_pseudolabels.py:55-101.  Local row i is global image idx0 + i: boards store global
 * indices and rank[] is indexed globally, F/pred/p_pred/probs locally — a shard continues on the
 * state handed over by the owner of the preceding index range. */
int gb_pseudolabel_scan(gb_ctx* ctx, void* state, const void* F, const void* T, float scale, int N,
                        int C, int k, int mode, int idx0, const int32_t* rank, int32_t* pred,
                        float* p_pred, float* probs, void* stream);

/* Counter that changes whenever the ctx re-allocates one of its scratch arenas (a call larger than any
 * before it).  Calls are otherwise allocation- and sync-free, so a sequence of them may be captured in a
 * CUDA graph; a captured graph must be re-captured when this value changes. */
uint64_t gb_workspace_generation(gb_ctx* ctx);

/* ---- prompt-tuning step glue on the device --------------------------------------
 * Cosine-logit cross-entropy of the reference's training loops and its gradient w.r.t. the text features
 * (methods/semi_supervised_learning/textual_prompt.py:93-109: normalise both sides, logits =
 * logit_scale.exp()·I·Tᵀ, nn.CrossEntropyLoss), in three launches and without a host round trip.
 *   imfn16 : fp16 [B,512] unit image features (gb_vit_forward's featn)     text : fp32 [C,512] UN-normalised
 *   labels : int32 [B]                      coef : fp32 [B] per-sample weights or NULL (= 1/B, mean reduction;
 *            the FPL variants' balance_param split, textual_fpl.py:123-165, is balance/|group| per sample)
 *   dtext  : fp32 [C,512] = d loss / d text (through the normalisation)   loss : fp32 [1] or NULL
 *   pred   : int32 [B] arg-max class per row or NULL                      deterministic (fixed summation order) */
int gb_ce_text_grad(gb_ctx* ctx, const void* imfn16, const float* text, const int32_t* labels,
                    const float* coef, float logit_scale_exp, int B, int C, float* dtext, float* loss,
                    int32_t* pred, void* stream);
/* The same loss with the image side trainable (VPT / UPT: methods/semi_supervised_learning/visual_prompt.py:122-135,
 * multimodal_prompt.py:103-121): image fp32 [B,512] and text fp32 [C,512] both UN-normalised; returns
 * dimage fp32 [B,512] = d loss / d image and / or dtext fp32 [C,512] = d loss / d text (either may be NULL, not
 * both), each through its side's normalisation.  Deterministic. */
int gb_ce_image_grad(gb_ctx* ctx, const float* image, const float* text, const int32_t* labels,
                     const float* coef, float logit_scale_exp, int B, int C, float* dimage, float* dtext,
                     float* loss, int32_t* pred, void* stream);
/* torch.optim.SGD step (dampening 0, no Nesterov): g += wd·p; buf = first ? g : mu·buf + g; p -= lr·(mu ? buf : g).
 * lr_dev != NULL: the learning rate is read from that device scalar instead of `lr` (a captured CUDA graph
 * of the step then keeps following the schedule). */
int gb_sgd_step(gb_ctx* ctx, float* param, const float* grad, float* momentum_buf, long long n, float lr,
                const float* lr_dev, float momentum, float weight_decay, int first_step, void* stream);
/* Learning rate of utils/schedulers.py:36-65 (WarmupCosineSchedule, cycles 0.5) at `step`. */
double gb_warmup_cosine_lr(double base_lr, int warmup_steps, int t_total, int step);

#ifdef __cplusplus
}
#endif
#endif /* GRIPB200_H_ */
