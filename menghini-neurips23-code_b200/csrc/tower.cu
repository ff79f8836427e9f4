// Tower-level entry points: the frozen CLIP ViT-B/32 image tower and causal text tower with the
// learnable prompt rows injected, forward and prompt-only backward.
//
// Reference: models/clip_encoders.py:43-90 (CustomTextEncoder.forward), :123-194
// (CustomVisionTransformer.forward) on top of clip.model.{VisionTransformer,Transformer,
// ResidualAttentionBlock}; gradients replace autograd through those modules w.r.t.
// TextPrefixModel.prefix / ImagePrefixModel.prefix (models/prompts_models.py:28,52) — the backbone is
// frozen, so only data gradients (dgrad) are ever formed.
//
// Layout: tokens of all samples form one flat row dimension M = samples·L (sample-major), fp16
// residual stream, fp32 LayerNorm statistics, fp32 accumulation in every contraction.  Each block is
//   GEMM(ln_1 folded, +bias) → attention → GEMM(+bias,+residual) → GEMM(ln_2 folded, +bias, QuickGELU)
//   → GEMM(+bias,+residual)
// = 5 launches; LayerNorm (as folded weights + epilogue row statistics), bias, activation and residual
// adds all live in GEMM epilogues.
#include <string.h>

#include "ctx.h"

// rowops.cu / attn.cu launchers
int gb_launch_layernorm(gb_ctx* c, const void* x, int ldx, const int32_t* row_idx, int in_row_mul,
                        const float* gamma, const float* beta, void* y, int ldy, int rows, int D,
                        int out_f32, cudaStream_t st);
int gb_launch_layernorm_bwd(gb_ctx* c, const void* dy, int lddy, const void* x, int ldx,
                            const int32_t* row_idx, int in_row_mul, const float* gamma, void* dx,
                            int lddx, int rows, int D, int accumulate, cudaStream_t st);
int gb_launch_im2col(gb_ctx* c, const void* img, int img_fmt, void* out, int B, cudaStream_t st);
int gb_launch_vit_assemble(gb_ctx* c, const void* patch, const float* cls, const float* pos,
                           const float* prefix, int P, const float* gamma, const float* beta,
                           void* x, int B, cudaStream_t st, float* stats);
int gb_launch_text_assemble(gb_ctx* c, const int32_t* ids, int ld_ids, const void* tok_emb,
                            const float* pos, const float* prefix, int P, void* x, int C, int Lt,
                            cudaStream_t st, float* stats);
int gb_launch_l2norm512(gb_ctx* c, const float* x, void* y16, float* y32, int rows, cudaStream_t st);
int gb_launch_prefix_grad(gb_ctx* c, const void* dx, int L, int S, int P, int D, const float* prefix,
                          const float* gamma, int ln_pre, float inv_scale, float* dprefix,
                          cudaStream_t st);
int gb_launch_scale_f32_to_f16(gb_ctx* c, const float* in, void* out, size_t n, float scale,
                               cudaStream_t st);
int gb_launch_eot_rows(gb_ctx* c, const int32_t* eot, int32_t* rows, int C, int Lt, cudaStream_t st);
int gb_launch_attn_fwd(gb_ctx* c, const void* qkv, void* out, int B, int L, int D, int causal,
                       cudaStream_t st);
int gb_launch_attn_bwd(gb_ctx* c, const void* qkv, const void* dout, void* dqkv, int B, int L,
                       int D, int causal, cudaStream_t st);

struct gb_tower {
  int width = 0, layers = 0, heads = 0, out_dim = 0;
  std::vector<gb_block_weights> blocks;
  gb_vit_weights vit;    // valid for the image tower
  gb_text_weights text;  // valid for the text tower
};
void gb_tower_free(gb_tower* t) { delete t; }

namespace {

// The fp16 gradient stream is loss-scaled by a power of two (exact) to keep small per-token
// gradients out of the fp16 subnormal range; prefix_grad_kernel divides it out in fp32.
constexpr float kGradScale = 256.0f;

struct Bump {
  uint8_t* base;
  size_t off = 0;
  explicit Bump(void* b) : base(reinterpret_cast<uint8_t*>(b)) {}
  void* take(size_t bytes) {
    void* p = base ? base + off : nullptr;
    off += (bytes + 255) & ~size_t(255);
    return p;
  }
};

inline size_t h2(size_t rows, size_t cols) { return rows * cols * 2; }

struct Tape {
  // per layer: x0 [M,D] | qkv [M,3D] | x1 [M,D] | f [M,4D];   then x_final [M,D]
  uint8_t* base;
  size_t M, D;
  size_t layer_bytes() const { return h2(M, 9 * D); }
  void* x0(int l) const { return base + l * layer_bytes(); }
  void* qkv(int l) const { return base + l * layer_bytes() + h2(M, D); }
  void* x1(int l) const { return base + l * layer_bytes() + h2(M, 4 * D); }
  void* f(int l) const { return base + l * layer_bytes() + h2(M, 5 * D); }
  void* x_final(int layers) const { return base + layers * layer_bytes(); }
};

// Runs `layers` residual attention blocks over x (fp16 [M,D]).  Without a tape x is updated in
// place; with a tape layer l reads tape.x0(l) and leaves its output in tape.x0(l+1) / x_final.
// When the block table carries LayerNorm-folded weights (s_qkv / s_fc non-null) ln_1 and ln_2 are not
// separate passes: the in-proj / c_fc GEMMs read the raw residual stream and apply the row statistics
// in their epilogue, and the two residual GEMMs emit the (Σ, Σ²) partials of the rows they write,
// which a tiny kernel turns into (μ·rstd, rstd) per row.
// st_fin: [M][2] (μ·rstd, rstd) of the current residual stream (valid for x on entry);
// st_part: [D/GB_STAT_SEG][M] float4 scratch for the shifted partial sums.
// first_rows != nullptr: the caller needs just row 0 of every sample from the last block (the image tower's
// CLS token, models/clip_encoders.py:189-192).  Everything after that block's attention — out-proj +
// residual, ln_2, c_fc, QuickGELU, c_proj + residual — is row-wise, so it runs on those S rows only, read in
// place through a row stride of L·D.  Without a tape the rows are left compact in first_rows [S, 6·D halves:
// x1 | x2 | 4·D of MLP scratch]; with a tape they are written to their own places in the tape (x1, f and
// x_final of the last block are then valid on the CLS rows only — run_blocks_bwd(cls_only) reads nothing
// else).  Per-row arithmetic is unchanged, so features and gradients are bit-identical to the dense
// evaluation; the reference computes and discards the other rows (5.6 % of the tower's FLOPs at L = 50).
int run_blocks(gb_ctx* c, const gb_tower* t, int S, int L, int causal, void* x, const Tape* tape,
               void* h, void* qkv_ws, void* a, void* g, float* st_fin, float* st_part, cudaStream_t st,
               void* first_rows = nullptr) {
  const int D = t->width, M = S * L;
  int rc;
  for (int l = 0; l < t->layers; ++l) {
    const gb_block_weights& w = t->blocks[l];
    void* x0 = tape ? tape->x0(l) : x;
    void* x1 = tape ? tape->x1(l) : x;
    void* x2 = tape ? (l + 1 < t->layers ? tape->x0(l + 1) : tape->x_final(t->layers)) : x;
    void* qkv = tape ? tape->qkv(l) : qkv_ws;
    const bool fold = w.s_qkv != nullptr && w.s_fc != nullptr;
    gb_gemm_ln emit; emit.stats_out = st_part;
    if (fold) {
      // layer 0 reads the statistics the assemble kernel finalized; later layers merge the partials the
      // previous c_proj GEMM emitted in their own epilogue
      gb_gemm_ln ln1; ln1.col_sum = w.s_qkv;
      if (l == 0) ln1.ln_stats = st_fin; else { ln1.ln_parts = st_part; ln1.nparts = D / GB_STAT_SEG; }
      if ((rc = gb_launch_gemm(c, x0, D, w.w_qkv, D, w.b_qkv, nullptr, 0, qkv, 3 * D, M, 3 * D, D, 0, 0, st, nullptr, &ln1))) return rc;
    } else {
      if ((rc = gb_launch_layernorm(c, x0, D, nullptr, 1, w.ln1_g, w.ln1_b, h, D, M, D, 0, st))) return rc;
      if ((rc = gb_launch_gemm(c, h, D, w.w_qkv, D, w.b_qkv, nullptr, 0, qkv, 3 * D, M, 3 * D, D, 0, 0, st))) return rc;
    }
    if ((rc = gb_launch_attn_fwd(c, qkv, a, S, L, D, causal, st))) return rc;
    if (first_rows && l + 1 == t->layers) {
      const int ldr = L * D;  // row 0 of sample s sits at row s·L of the [M, D] buffers
      // compact scratch without a tape; the rows' own places in the tape (and in g) with one
      void* xc1 = tape ? x1 : first_rows;
      void* xc2 = tape ? x2 : reinterpret_cast<uint8_t*>(first_rows) + h2(S, D);
      void* gc = tape ? g : reinterpret_cast<uint8_t*>(first_rows) + 2 * h2(S, D);
      const int ld1 = tape ? ldr : D, ldg = tape ? 4 * ldr : 4 * D;
      if ((rc = gb_launch_gemm(c, a, ldr, w.w_o, D, w.b_o, x0, ldr, xc1, ld1, S, D, D, 0, 0, st, nullptr, fold ? &emit : nullptr))) return rc;
      if (fold) {
        gb_gemm_ln ln2; ln2.ln_parts = st_part; ln2.nparts = D / GB_STAT_SEG; ln2.col_sum = w.s_fc;
        if ((rc = gb_launch_gemm(c, xc1, ld1, w.w_fc, D, w.b_fc, nullptr, 0, gc, ldg, S, 4 * D, D, 1, 0, st,
                                 tape ? tape->f(l) : nullptr, &ln2))) return rc;
      } else {
        if ((rc = gb_launch_layernorm(c, xc1, ld1, nullptr, 1, w.ln2_g, w.ln2_b, h, D, S, D, 0, st))) return rc;
        if ((rc = gb_launch_gemm(c, h, D, w.w_fc, D, w.b_fc, nullptr, 0, gc, ldg, S, 4 * D, D, 1, 0, st,
                                 tape ? tape->f(l) : nullptr))) return rc;
      }
      if ((rc = gb_launch_gemm(c, gc, ldg, w.w_proj, 4 * D, w.b_proj, xc1, ld1, xc2, ld1, S, D, 4 * D, 0, 0, st))) return rc;
      break;
    }
    if ((rc = gb_launch_gemm(c, a, D, w.w_o, D, w.b_o, x0, D, x1, D, M, D, D, 0, 0, st, nullptr, fold ? &emit : nullptr))) return rc;
    if (fold) {
      gb_gemm_ln ln2; ln2.ln_parts = st_part; ln2.nparts = D / GB_STAT_SEG; ln2.col_sum = w.s_fc;
      if ((rc = gb_launch_gemm(c, x1, D, w.w_fc, D, w.b_fc, nullptr, 0, g, 4 * D, M, 4 * D, D, 1, 0, st,
                               tape ? tape->f(l) : nullptr, &ln2))) return rc;
    } else {
      if ((rc = gb_launch_layernorm(c, x1, D, nullptr, 1, w.ln2_g, w.ln2_b, h, D, M, D, 0, st))) return rc;
      if ((rc = gb_launch_gemm(c, h, D, w.w_fc, D, w.b_fc, nullptr, 0, g, 4 * D, M, 4 * D, D, 1, 0, st,
                               tape ? tape->f(l) : nullptr))) return rc;
    }
    const bool more = fold && l + 1 < t->layers;
    if ((rc = gb_launch_gemm(c, g, 4 * D, w.w_proj, 4 * D, w.b_proj, x1, D, x2, D, M, D, 4 * D, 0, 0, st, nullptr, more ? &emit : nullptr))) return rc;
  }
  return GB_OK;
}

// Reverse pass through the blocks: dx (fp16 [M,D], loss-scaled) is d loss / d x_final on entry and
// d loss / d x0(layer 0) on exit.
// cls_only: d loss / d x_final is non-zero on row 0 of every sample only (and zero elsewhere in dx), and the
// forward left x1 / f of the last block valid on those rows only (run_blocks with first_rows + tape): the last
// block's MLP and out-proj backward run on the S CLS rows through a row stride of L·D; from its attention
// backward on everything is dense (every key and value contributes to the CLS query).
int run_blocks_bwd(gb_ctx* c, const gb_tower* t, int S, int L, int causal, void* dx,
                   const Tape& tape, void* dh, void* dqkv, void* dg, cudaStream_t st, bool cls_only = false) {
  const int D = t->width, M = S * L;
  int rc;
  for (int l = t->layers - 1; l >= 0; --l) {
    const gb_block_weights& w = t->blocks[l];
    if (!w.w_qkv_t || !w.w_o_t || !w.w_fc_t || !w.w_proj_t)
      return gb_fail(c, GB_ERR_STATE, "backward: transposed weights were not provided");
    if (cls_only && l + 1 == t->layers) {
      const int ldr = L * D;
      if ((rc = gb_launch_gemm(c, dx, ldr, w.w_proj_t, D, nullptr, nullptr, 0, dg, 4 * ldr, S, 4 * D, D, 2, 0, st, tape.f(l)))) return rc;
      GB_CUDA(c, cudaMemsetAsync(dh, 0, h2(M, D), st));  // the attention backward reads dO of every row
      if ((rc = gb_launch_gemm(c, dg, 4 * ldr, w.w_fc_t, 4 * D, nullptr, nullptr, 0, dh, ldr, S, D, 4 * D, 0, 0, st))) return rc;
      if ((rc = gb_launch_layernorm_bwd(c, dh, ldr, tape.x1(l), ldr, nullptr, 1, w.ln2_g, dx, ldr, S, D, 1, st))) return rc;
      if ((rc = gb_launch_gemm(c, dx, ldr, w.w_o_t, D, nullptr, nullptr, 0, dh, ldr, S, D, D, 0, 0, st))) return rc;
      if ((rc = gb_launch_attn_bwd(c, tape.qkv(l), dh, dqkv, S, L, D, causal, st))) return rc;
      if ((rc = gb_launch_gemm(c, dqkv, 3 * D, w.w_qkv_t, 3 * D, nullptr, nullptr, 0, dh, D, M, D, 3 * D, 0, 0, st))) return rc;
      if ((rc = gb_launch_layernorm_bwd(c, dh, D, tape.x0(l), D, nullptr, 1, w.ln1_g, dx, D, M, D, 1, st))) return rc;
      continue;
    }
    // MLP branch: x2 = x1 + c_proj(QuickGELU(c_fc(ln_2(x1))))
    if ((rc = gb_launch_gemm(c, dx, D, w.w_proj_t, D, nullptr, nullptr, 0, dg, 4 * D, M, 4 * D, D, 2, 0, st, tape.f(l)))) return rc;
    if ((rc = gb_launch_gemm(c, dg, 4 * D, w.w_fc_t, 4 * D, nullptr, nullptr, 0, dh, D, M, D, 4 * D, 0, 0, st))) return rc;
    if ((rc = gb_launch_layernorm_bwd(c, dh, D, tape.x1(l), D, nullptr, 1, w.ln2_g, dx, D, M, D, 1, st))) return rc;
    // attention branch: x1 = x0 + out_proj(attn(in_proj(ln_1(x0))))
    if ((rc = gb_launch_gemm(c, dx, D, w.w_o_t, D, nullptr, nullptr, 0, dh, D, M, D, D, 0, 0, st))) return rc;
    if ((rc = gb_launch_attn_bwd(c, tape.qkv(l), dh, dqkv, S, L, D, causal, st))) return rc;
    if ((rc = gb_launch_gemm(c, dqkv, 3 * D, w.w_qkv_t, 3 * D, nullptr, nullptr, 0, dh, D, M, D, 3 * D, 0, 0, st))) return rc;
    if ((rc = gb_launch_layernorm_bwd(c, dh, D, tape.x0(l), D, nullptr, 1, w.ln1_g, dx, D, M, D, 1, st))) return rc;
  }
  return GB_OK;
}

int copy_blocks(gb_ctx* c, gb_tower* t, const gb_block_weights* blocks, int layers) {
  if (!blocks || layers <= 0 || layers > 64) return gb_fail(c, GB_ERR_ARG, "set_weights: bad block table");
  t->blocks.assign(blocks, blocks + layers);
  for (const gb_block_weights& w : t->blocks) {
    if (!w.ln1_g || !w.ln1_b || !w.w_qkv || !w.b_qkv || !w.w_o || !w.b_o || !w.ln2_g || !w.ln2_b ||
        !w.w_fc || !w.b_fc || !w.w_proj || !w.b_proj)
      return gb_fail(c, GB_ERR_ARG, "set_weights: null weight pointer in block table");
  }
  return GB_OK;
}

}  // namespace

extern "C" size_t gb_tape_bytes(int samples, int L, int D, int layers) {
  if (samples <= 0 || L <= 0 || D <= 0 || layers <= 0) return 0;
  return ((size_t)layers * 9 + 1) * (size_t)samples * L * D * 2;
}

extern "C" int gb_vit_set_weights(gb_ctx* c, const gb_vit_weights* w) {
  if (!c) return GB_ERR_ARG;
  gb_dev_guard dev_guard(c);
  if (!w || w->width != 768 || w->out_dim != 512 || w->heads * 64 != w->width)
    return gb_fail(c, GB_ERR_ARG, "vit_set_weights: only ViT-B/32 geometry (width 768, 12 heads, out 512)");
  if (!w->conv_w || !w->cls || !w->pos || !w->ln_pre_g || !w->ln_pre_b || !w->ln_post_g ||
      !w->ln_post_b || !w->proj_t)
    return gb_fail(c, GB_ERR_ARG, "vit_set_weights: null weight pointer");
  gb_tower* t = new gb_tower();
  int rc = copy_blocks(c, t, w->blocks, w->layers);
  if (rc) { delete t; return rc; }
  t->width = w->width; t->layers = w->layers; t->heads = w->heads; t->out_dim = w->out_dim;
  t->vit = *w;
  t->vit.blocks = t->blocks.data();
  delete c->vit;
  c->vit = t;
  return GB_OK;
}

extern "C" int gb_text_set_weights(gb_ctx* c, const gb_text_weights* w) {
  if (!c) return GB_ERR_ARG;
  gb_dev_guard dev_guard(c);
  if (!w || w->width != 512 || w->out_dim != 512 || w->heads * 64 != w->width || w->ctx_len > 96)
    return gb_fail(c, GB_ERR_ARG, "text_set_weights: only the ViT-B/32 text geometry (width 512, 8 heads, out 512)");
  if (!w->tok_emb || !w->pos || !w->ln_final_g || !w->ln_final_b || !w->proj_t)
    return gb_fail(c, GB_ERR_ARG, "text_set_weights: null weight pointer");
  gb_tower* t = new gb_tower();
  int rc = copy_blocks(c, t, w->blocks, w->layers);
  if (rc) { delete t; return rc; }
  t->width = w->width; t->layers = w->layers; t->heads = w->heads; t->out_dim = w->out_dim;
  t->text = *w;
  t->text.blocks = t->blocks.data();
  delete c->text;
  c->text = t;
  return GB_OK;
}

extern "C" int gb_vit_forward(gb_ctx* c, const void* img, int img_f32, const float* prefix, int B,
                              int P, float* feat, void* featn, void* tape_mem, void* stream) {
  if (!c) return GB_ERR_ARG;
  gb_dev_guard dev_guard(c);
  gb_tower* t = c->vit;
  if (!t) return gb_fail(c, GB_ERR_STATE, "vit_forward: weights not set");
  if (B <= 0) return GB_OK;
  if (!img || P < 0 || P > 46 || (P > 0 && !prefix) || (!feat && !featn))
    return gb_fail(c, GB_ERR_ARG, "vit_forward: bad arguments (B=%d P=%d)", B, P);
  cudaStream_t st = (cudaStream_t)stream;
  const int D = 768, L = 50 + P;
  const size_t M = (size_t)B * L;
  // workspace: x | h | qkv | a | g (im2col aliases g, the patch embeddings alias a) | cls | feat
  size_t need;
  {
    Bump b(nullptr);
    b.take(h2(M, D)); b.take(h2(M, D)); b.take(h2(M, 3 * D)); b.take(h2(M, D));
    b.take(h2(M > (size_t)B * 49 ? M : (size_t)B * 49, 4 * D)); b.take(h2(B, D)); b.take((size_t)B * 512 * 4);
    b.take(M * 8); b.take(M * (D / GB_STAT_SEG) * 16); b.take(h2(B, 6 * D));
    need = b.off;
  }
  int rc = gb_ws_reserve(c, gb_ctx::kWsVit, need);
  if (rc) return rc;
  Bump b(c->ws[gb_ctx::kWsVit]);
  void* x_ws = b.take(h2(M, D));
  void* h = b.take(h2(M, D));
  void* qkv = b.take(h2(M, 3 * D));
  void* a = b.take(h2(M, D));
  void* g = b.take(h2(M > (size_t)B * 49 ? M : (size_t)B * 49, 4 * D));
  void* cls_ln = b.take(h2(B, D));
  float* feat_ws = reinterpret_cast<float*>(b.take((size_t)B * 512 * 4));
  float* st_a = reinterpret_cast<float*>(b.take(M * 8));                 // (μ·rstd, rstd) per row
  float* st_b = reinterpret_cast<float*>(b.take(M * (D / GB_STAT_SEG) * 16));    // shifted partials per GB_STAT_SEG columns
  void* cls_rows = b.take(h2(B, 6 * D));                                 // last block on the CLS rows only
  Tape tape{reinterpret_cast<uint8_t*>(tape_mem), M, (size_t)D};
  void* x = tape_mem ? tape.x0(0) : x_ws;
  const gb_vit_weights& w = t->vit;
  // conv1 as im2col + GEMM (kernel == stride ⇒ im2col is a permutation): models/clip_encoders.py:131-133
  if ((rc = gb_launch_im2col(c, img, img_f32, g, B, st))) return rc;
  if ((rc = gb_launch_gemm(c, g, 3072, w.conv_w, 3072, nullptr, nullptr, 0, a, D, B * 49, D, 3072, 0, 0, st))) return rc;
  // CLS + pos-emb, prefix rows, ln_pre: :135-157
  if ((rc = gb_launch_vit_assemble(c, a, w.cls, w.pos, prefix, P, w.ln_pre_g, w.ln_pre_b, x, B, st, st_a))) return rc;
  // only the CLS rows of the last block are evaluated past its attention (see run_blocks)
  if ((rc = run_blocks(c, t, B, L, 0, x, tape_mem ? &tape : nullptr, h, qkv, a, g, st_a, st_b, st, cls_rows))) return rc;
  // ln_post(x[:,0,:]) @ proj: :189-192
  if (tape_mem) {
    if ((rc = gb_launch_layernorm(c, tape.x_final(t->layers), D, nullptr, L, w.ln_post_g, w.ln_post_b, cls_ln, D, B, D, 0, st))) return rc;
  } else {
    const uint8_t* xc2 = reinterpret_cast<const uint8_t*>(cls_rows) + h2(B, D);
    if ((rc = gb_launch_layernorm(c, xc2, D, nullptr, 1, w.ln_post_g, w.ln_post_b, cls_ln, D, B, D, 0, st))) return rc;
  }
  float* fo = feat ? feat : feat_ws;
  if ((rc = gb_launch_gemm(c, cls_ln, D, w.proj_t, D, nullptr, nullptr, 0, fo, 512, B, 512, D, 0, 1, st))) return rc;
  if (featn && (rc = gb_launch_l2norm512(c, fo, featn, nullptr, B, st))) return rc;
  return GB_OK;
}

extern "C" int gb_vit_backward_prefix(gb_ctx* c, const float* dfeat, const float* prefix, int B,
                                      int P, const void* tape_mem, float* dprefix, void* stream) {
  if (!c) return GB_ERR_ARG;
  gb_dev_guard dev_guard(c);
  gb_tower* t = c->vit;
  if (!t) return gb_fail(c, GB_ERR_STATE, "vit_backward_prefix: weights not set");
  if (!dfeat || !prefix || !tape_mem || !dprefix || B <= 0 || P <= 0 || P > 46)
    return gb_fail(c, GB_ERR_ARG, "vit_backward_prefix: bad arguments");
  if (!t->vit.proj) return gb_fail(c, GB_ERR_STATE, "vit_backward_prefix: proj (untransposed) not provided");
  cudaStream_t st = (cudaStream_t)stream;
  const int D = 768, L = 50 + P;
  const size_t M = (size_t)B * L;
  size_t need;
  {
    Bump b(nullptr);
    b.take(h2(M, D)); b.take(h2(M, D)); b.take(h2(M, 3 * D)); b.take(h2(M, 4 * D));
    b.take(h2(B, 512)); b.take(h2(B, D));
    need = b.off;
  }
  int rc = gb_ws_reserve(c, gb_ctx::kWsVit, need);
  if (rc) return rc;
  Bump b(c->ws[gb_ctx::kWsVit]);
  void* dx = b.take(h2(M, D));
  void* dh = b.take(h2(M, D));
  void* dqkv = b.take(h2(M, 3 * D));
  void* dg = b.take(h2(M, 4 * D));
  void* d16 = b.take(h2(B, 512));
  void* dcls = b.take(h2(B, D));
  Tape tape{reinterpret_cast<uint8_t*>(const_cast<void*>(tape_mem)), M, (size_t)D};
  const gb_vit_weights& w = t->vit;
  if ((rc = gb_launch_scale_f32_to_f16(c, dfeat, d16, (size_t)B * 512, kGradScale, st))) return rc;
  // feat = ln_post(cls) @ proj  →  d ln_post-out = dfeat @ projᵀ
  if ((rc = gb_launch_gemm(c, d16, 512, w.proj, 512, nullptr, nullptr, 0, dcls, D, B, D, 512, 0, 0, st))) return rc;
  GB_CUDA(c, cudaMemsetAsync(dx, 0, h2(M, D), st));
  if ((rc = gb_launch_layernorm_bwd(c, dcls, D, tape.x_final(t->layers), D, nullptr, L, w.ln_post_g, dx, D, B, D, 0, st))) return rc;
  if ((rc = run_blocks_bwd(c, t, B, L, 0, dx, tape, dh, dqkv, dg, st, /*cls_only=*/true))) return rc;
  return gb_launch_prefix_grad(c, dx, L, B, P, D, prefix, w.ln_pre_g, 1, 1.0f / kGradScale, dprefix, st);
}

extern "C" int gb_text_forward(gb_ctx* c, const int32_t* ids, int ld_ids, const int32_t* eot,
                               const float* prefix, int C, int P, int Lt, float* feat, void* featn,
                               void* tape_mem, void* stream) {
  if (!c) return GB_ERR_ARG;
  gb_dev_guard dev_guard(c);
  gb_tower* t = c->text;
  if (!t) return gb_fail(c, GB_ERR_STATE, "text_forward: weights not set");
  if (C <= 0) return GB_OK;
  const gb_text_weights& w = t->text;
  if (!ids || !eot || Lt < 1 || Lt > w.ctx_len || ld_ids < Lt || P < 0 || P + 2 > Lt ||
      (P > 0 && !prefix) || (!feat && !featn))
    return gb_fail(c, GB_ERR_ARG, "text_forward: bad arguments (C=%d P=%d Lt=%d)", C, P, Lt);
  cudaStream_t st = (cudaStream_t)stream;
  const int D = 512, L = Lt;
  const size_t M = (size_t)C * L;
  size_t need;
  {
    Bump b(nullptr);
    b.take(h2(M, D)); b.take(h2(M, D)); b.take(h2(M, 3 * D)); b.take(h2(M, D)); b.take(h2(M, 4 * D));
    b.take(h2(C, D)); b.take((size_t)C * 512 * 4); b.take((size_t)C * 4);
    b.take(M * 8); b.take(M * (D / GB_STAT_SEG) * 16);
    need = b.off;
  }
  int rc = gb_ws_reserve(c, gb_ctx::kWsText, need);
  if (rc) return rc;
  Bump b(c->ws[gb_ctx::kWsText]);
  void* x_ws = b.take(h2(M, D));
  void* h = b.take(h2(M, D));
  void* qkv = b.take(h2(M, 3 * D));
  void* a = b.take(h2(M, D));
  void* g = b.take(h2(M, 4 * D));
  void* eot_ln = b.take(h2(C, D));
  float* feat_ws = reinterpret_cast<float*>(b.take((size_t)C * 512 * 4));
  int32_t* rows = reinterpret_cast<int32_t*>(b.take((size_t)C * 4));
  float* st_a = reinterpret_cast<float*>(b.take(M * 8));                 // (μ·rstd, rstd) per row
  float* st_b = reinterpret_cast<float*>(b.take(M * (D / GB_STAT_SEG) * 16));    // shifted partials per GB_STAT_SEG columns
  Tape tape{reinterpret_cast<uint8_t*>(tape_mem), M, (size_t)D};
  void* x = tape_mem ? tape.x0(0) : x_ws;
  // token embedding, prefix overwrite of rows 1..P, + positional embedding: models/clip_encoders.py:63-74
  if ((rc = gb_launch_text_assemble(c, ids, ld_ids, w.tok_emb, w.pos, prefix, P, x, C, L, st, st_a))) return rc;
  if ((rc = run_blocks(c, t, C, L, 1, x, tape_mem ? &tape : nullptr, h, qkv, a, g, st_a, st_b, st))) return rc;
  const void* xf = tape_mem ? tape.x_final(t->layers) : x;
  // ln_final, EOT-row gather, @ text_projection: :85-89
  if ((rc = gb_launch_eot_rows(c, eot, rows, C, L, st))) return rc;
  if ((rc = gb_launch_layernorm(c, xf, D, rows, 1, w.ln_final_g, w.ln_final_b, eot_ln, D, C, D, 0, st))) return rc;
  float* fo = feat ? feat : feat_ws;
  if ((rc = gb_launch_gemm(c, eot_ln, D, w.proj_t, D, nullptr, nullptr, 0, fo, 512, C, 512, D, 0, 1, st))) return rc;
  if (featn && (rc = gb_launch_l2norm512(c, fo, featn, nullptr, C, st))) return rc;
  return GB_OK;
}

extern "C" int gb_text_backward_prefix(gb_ctx* c, const float* dfeat, const int32_t* eot, int C,
                                       int P, int Lt, const void* tape_mem, float* dprefix,
                                       void* stream) {
  if (!c) return GB_ERR_ARG;
  gb_dev_guard dev_guard(c);
  gb_tower* t = c->text;
  if (!t) return gb_fail(c, GB_ERR_STATE, "text_backward_prefix: weights not set");
  const gb_text_weights& w = t->text;
  if (!dfeat || !eot || !tape_mem || !dprefix || C <= 0 || P <= 0 || Lt < P + 2 || Lt > w.ctx_len)
    return gb_fail(c, GB_ERR_ARG, "text_backward_prefix: bad arguments");
  if (!w.proj) return gb_fail(c, GB_ERR_STATE, "text_backward_prefix: text_projection (untransposed) not provided");
  cudaStream_t st = (cudaStream_t)stream;
  const int D = 512, L = Lt;
  const size_t M = (size_t)C * L;
  size_t need;
  {
    Bump b(nullptr);
    b.take(h2(M, D)); b.take(h2(M, D)); b.take(h2(M, 3 * D)); b.take(h2(M, 4 * D));
    b.take(h2(C, 512)); b.take(h2(C, D)); b.take((size_t)C * 4);
    need = b.off;
  }
  int rc = gb_ws_reserve(c, gb_ctx::kWsText, need);
  if (rc) return rc;
  Bump b(c->ws[gb_ctx::kWsText]);
  void* dx = b.take(h2(M, D));
  void* dh = b.take(h2(M, D));
  void* dqkv = b.take(h2(M, 3 * D));
  void* dg = b.take(h2(M, 4 * D));
  void* d16 = b.take(h2(C, 512));
  void* deot = b.take(h2(C, D));
  int32_t* rows = reinterpret_cast<int32_t*>(b.take((size_t)C * 4));
  Tape tape{reinterpret_cast<uint8_t*>(const_cast<void*>(tape_mem)), M, (size_t)D};
  if ((rc = gb_launch_scale_f32_to_f16(c, dfeat, d16, (size_t)C * 512, kGradScale, st))) return rc;
  if ((rc = gb_launch_gemm(c, d16, 512, w.proj, 512, nullptr, nullptr, 0, deot, D, C, D, 512, 0, 0, st))) return rc;
  if ((rc = gb_launch_eot_rows(c, eot, rows, C, L, st))) return rc;
  GB_CUDA(c, cudaMemsetAsync(dx, 0, h2(M, D), st));
  if ((rc = gb_launch_layernorm_bwd(c, deot, D, tape.x_final(t->layers), D, rows, 1, w.ln_final_g, dx, D, C, D, 0, st))) return rc;
  if ((rc = run_blocks_bwd(c, t, C, L, 1, dx, tape, dh, dqkv, dg, st))) return rc;
  return gb_launch_prefix_grad(c, dx, L, C, P, D, nullptr, nullptr, 0, 1.0f / kGradScale, dprefix, st);
}
