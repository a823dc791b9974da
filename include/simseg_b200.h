/* libsimseg_b200 — C ABI of the B200-native SimSeg hot path.
 *
 * The reference (muyangyi/SimSeg) is 100% Python; it has no FFI of its own.  Its boundary for
 * this path is the nn.Module contract of simseg/models/pipelines/clip.py:13-229.  This header
 * is the C ABI we put UNDER that contract: each entry point names the reference call it
 * replaces (paths relative to the reference root).  The Python mirror of the reference
 * interface (package simseg_b200) binds these with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *  - every function returns 0 on success, <0 on error (simseg_last_error() gives the text,
 *    thread-local);
 *  - tensors are raw DEVICE pointers + explicit sizes; row-major unless stated; the caller
 *    owns every buffer including workspaces (no hidden allocation, no host sync);
 *  - `stream` is a cudaStream_t passed as void*; work is enqueued, never synchronised;
 *  - dtype enum: SIMSEG_F32 = 0, SIMSEG_BF16 = 1;
 *  - one simseg_ctx per process/GPU; calls on one ctx are not thread-safe.
 */
#ifndef SIMSEG_B200_H_
#define SIMSEG_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SIMSEG_OK 0
#define SIMSEG_ERR_INVALID (-1)
#define SIMSEG_ERR_CUDA (-2)
#define SIMSEG_ERR_UNSUPPORTED (-3)

#define SIMSEG_F32 0
#define SIMSEG_BF16 1

typedef struct simseg_ctx simseg_ctx;

/* ---- context ------------------------------------------------------------------------------ */
int simseg_version(void);
const char* simseg_last_error(void);
int simseg_ctx_create(int device, simseg_ctx** out);
int simseg_ctx_destroy(simseg_ctx* ctx);
/* number of kernels launched through this ctx since creation / last reset */
int64_t simseg_ctx_launch_count(simseg_ctx* ctx, int reset);

/* ---- GEMM engine (tcgen05 + TMA) ---------------------------------------------------------- */
/* Replaces every nn.Linear on the path: timm Block qkv/proj/fc1/fc2 driven by
 * simseg/models/backbones/mml/vit_builder.py:18, HF BertLayer dense layers
 * (huggingface_builder.py:16-17), SimpleProjection (components/projection.py:45-46), the
 * patch-embed conv as GEMM (vit_builder.py:14), and their dgrad / wgrad in backward.
 *
 *   D[M,N] = epilogue( sum_k A[m,k] * B[n,k] )       bf16 operands, fp32 accumulation in TMEM
 *
 * a_major / b_major: 0 = K-major  (A stored [M,K] row-major, B stored [N,K] row-major, i.e. x @ W^T)
 *                    1 = MN-major (A stored [K,M] row-major, B stored [K,N] row-major)
 * lda / ldb / ldd / ld_aux / ld_res: leading dimension in elements of the respective storage. */
#define SIMSEG_EPI_NONE 0          /* D = acc (+bias)                                     */
#define SIMSEG_EPI_BIAS_GELU 1     /* aux(bf16, optional) = acc+bias ; D = gelu_erf(acc+bias) */
#define SIMSEG_EPI_BIAS_RESIDUAL 2 /* D = acc + bias + residual(f32 or bf16, dtype res_dtype) */
#define SIMSEG_EPI_DGELU 3         /* D = acc * gelu_erf'(aux[m,n])   (aux = saved pre-activation) ; aux2 = gelu_erf(aux) */
#define SIMSEG_EPI_ROWSCALE 4      /* D = acc * row_scale[m]                              */

typedef struct simseg_gemm_args {
  const void* a;
  const void* b;
  void* d;
  int64_t M, N, K;
  int64_t lda, ldb, ldd;
  int32_t a_major, b_major;
  int32_t in_dtype;       /* SIMSEG_BF16 (tcgen05 kind::f16) or SIMSEG_F32 (kind::tf32) */
  int32_t out_dtype;      /* SIMSEG_F32 or SIMSEG_BF16 */
  int32_t epilogue;       /* SIMSEG_EPI_* */
  int32_t accumulate;     /* !=0: D += result (wgrad accumulation across micro-batches; out f32 only) */
  const float* bias;      /* [N] fp32 or NULL */
  const void* residual;   /* [M,N] or NULL */
  int64_t ld_res;
  int32_t res_dtype;
  void* aux;              /* [M,N] bf16: out for BIAS_GELU (may be NULL), in for DGELU */
  int64_t ld_aux;
  const float* row_scale; /* [M] for ROWSCALE */
  float* col_sum;         /* optional [N] fp32: atomically accumulates column sums of the fp32
                             epilogue result (bias gradients) ; NULL = off */
  int32_t tile_n;         /* 0 = auto; else 128/192/256 */
  int32_t reserved;
  void* aux2;             /* DGELU only, optional [M,N] bf16 out: gelu_erf(aux) (the activation the following
                             wgrad needs), produced by the same epilogue instead of a separate recompute pass */
  int64_t ld_aux2;
} simseg_gemm_args;

int simseg_gemm(simseg_ctx* ctx, const simseg_gemm_args* args, void* stream);

/* ---- elementwise / normalisation ----------------------------------------------------------- */
/* fp32 -> bf16 cast, optionally also writing the transpose ([rows,cols] -> [cols,rows]). */
int simseg_cast_bf16(simseg_ctx* ctx, const float* src, void* dst, void* dst_t, int64_t rows, int64_t cols,
                     void* stream);
/* The same cast for many weights in one launch (the per-optimizer-step bf16 refresh of every Linear weight).  `items` is a
 * DEVICE array; item i covers blocks [first_block, first_block + ceil(rows/32)*ceil(cols/128)) of the 1-D grid of
 * `total_blocks` blocks (first_block ascending).  dst rows have pitch `ld` (>= cols), dst_t rows pitch `ld_t` (>= rows) so
 * several weights can land side by side in one packed matrix (BERT q/k/v); dst or dst_t may be NULL. */
typedef struct simseg_cast_item {
  const float* src;
  void* dst;
  void* dst_t;
  int64_t rows, cols, ld, ld_t, first_block;
} simseg_cast_item;
int simseg_cast_bf16_multi(simseg_ctx* ctx, const simseg_cast_item* items, int n_items, int64_t total_blocks, void* stream);
/* bf16/f32 column sums: out[n] (+)= sum_m x[m,n]   (bias gradients) */
int simseg_colsum(simseg_ctx* ctx, const void* x, int dtype, int64_t M, int64_t N, int64_t ldx, float* out,
                  int accumulate, void* stream);
/* a = gelu_erf(h), bf16 -> bf16 (timm Mlp / HF BertIntermediate activation as a standalone pass; the towers do not use it:
 * forward applies GELU in the fc1 GEMM epilogue, backward gets gelu(h) kept from the forward or re-emitted by the dGELU epilogue) */
int simseg_gelu_fwd(simseg_ctx* ctx, const void* h, void* a, int64_t n, void* stream);

/* LayerNorm over the last dim (timm LayerNorm eps 1e-6, HF BERT LayerNorm eps 1e-12).
 * x: [M,D] f32 or bf16; writes y_bf16 and/or y_f32 (either may be NULL), mean/rstd [M] (may be NULL). */
int simseg_layernorm_fwd(simseg_ctx* ctx, const void* x, int x_dtype, const float* gamma, const float* beta,
                         float eps, int64_t M, int D, void* y_bf16, float* y_f32, float* mean, float* rstd,
                         void* stream);
/* Residual add fused with the LayerNorm that follows it (timm Block: x + attn(..) -> norm2, x + mlp(..) -> next norm1;
 * HF BertSelfOutput / BertOutput: LayerNorm(dense(..) + input)):  s = x (f32) + add (bf16 output of the Linear, exactly
 * what the reference's autocast produces) ; sum_out (f32, optional) = s ; y = LayerNorm(s). */
int simseg_add_layernorm_fwd(simseg_ctx* ctx, const float* x, const void* add_bf16, const float* gamma, const float* beta,
                             float eps, int64_t M, int D, float* sum_out, void* y_bf16, float* y_f32, float* mean, float* rstd,
                             void* stream);
/* dx[M,D] (f32) = LN backward ; if dx_accumulate, dx += (residual-gradient accumulation);
 * dgamma/dbeta [D] fp32 are ACCUMULATED atomically (zero them first).
 * dy: f32 or bf16; optional second upstream gradient dy2 (f32, may be NULL) is added to dy first.
 * dx_colsum [D] (may be NULL): ACCUMULATES the column sums of the final dx — the bias gradient of the
 * residual-branch Linear that produced x's summand. */
int simseg_layernorm_bwd(simseg_ctx* ctx, const void* dy, int dy_dtype, const float* dy2, const void* x,
                         int x_dtype, const float* gamma, const float* mean, const float* rstd, int64_t M, int D,
                         float* dx, int dx_accumulate, void* dx_bf16, float* dgamma, float* dbeta, float* dx_colsum,
                         void* stream);

/* ---- attention (timm Attention.forward; HF BertSelfAttention) ------------------------------- */
/* q,k,v: bf16 with element strides (batch, token, head); head_dim = 64 contiguous.
 * ViT packs qkv as [B,S,3,H,64] (vit_builder.py:18 -> timm Attention); BERT has three [B,T,H*64].
 * key_len: optional int32 [B] — keys >= key_len[b] are masked (BERT additive mask from a
 * left-aligned attention_mask); NULL = no mask.
 * out: bf16 [B,S,H*64]; lse: fp32 [B,H,S] (log-sum-exp of scaled scores, saved for backward). */
int simseg_attention_fwd(simseg_ctx* ctx, const void* q, const void* k, const void* v, int64_t stride_b,
                         int64_t stride_s, int64_t stride_h, int B, int H, int S, const int32_t* key_len,
                         float scale, void* out, float* lse, void* stream);
/* dq,dk,dv written with the same strides as q,k,v (so a packed dqkv buffer works). */
int simseg_attention_bwd(simseg_ctx* ctx, const void* q, const void* k, const void* v, const void* out,
                         const void* dout, const float* lse, int64_t stride_b, int64_t stride_s, int64_t stride_h,
                         int B, int H, int S, const int32_t* key_len, float scale, void* dq, void* dk, void* dv,
                         void* stream);

/* ---- train-mode dropout of the BERT tower ------------------------------------------------------
 * HF BertModel under model.train() (huggingface_builder.py:16-17 wraps it; hidden_dropout_prob =
 * attention_probs_dropout_prob = 0.1 in bert-base-uncased): dropout after the embedding LayerNorm, on the attention
 * probabilities, and on the two dense outputs before their residual add + LayerNorm.
 * A keep bit is a pure function of (seed, step, site, element) — Philox4x32-10, csrc/philox.cuh — so backward regenerates
 * the forward's mask.  drop_rng is a DEVICE pointer to {seed, step}: kernels read it at run time, a replayed CUDA graph
 * draws new masks once the host (or a captured add) bumps step.  drop_site separates the dropout layers of one step.
 * Kept values are scaled by 1 / (1 - drop_p); drop_p = 0 reduces every call to its plain counterpart. */
/* add_bf16 != NULL: y = LayerNorm(x + dropout(add))            (BertSelfOutput / BertOutput), sum_out = x + dropout(add)
 * add_bf16 == NULL: y = dropout(LayerNorm(x))                  (BertEmbeddings) */
int simseg_layernorm_fwd_dropout(simseg_ctx* ctx, const float* x, const void* add_bf16, const float* gamma,
                                 const float* beta, float eps, int64_t M, int D, float* sum_out, void* y_bf16, float* y_f32,
                                 float* mean, float* rstd, float drop_p, const uint64_t* drop_rng, uint32_t drop_site,
                                 void* stream);
/* simseg_layernorm_bwd with the matching mask.  drop_mode 1 (forward had add_bf16): dx_bf16 and dx_colsum — the gradient of
 * the dropped summand and of its Linear's bias — carry the mask, dx (residual path) does not.  drop_mode 2 (forward
 * dropped the outputs): dy (+ dy2) is masked before the LayerNorm backward. */
int simseg_layernorm_bwd_dropout(simseg_ctx* ctx, const void* dy, int dy_dtype, const float* dy2, const float* x,
                                 const float* gamma, const float* mean, const float* rstd, int64_t M, int D, float* dx,
                                 int dx_accumulate, void* dx_bf16, float* dgamma, float* dbeta, float* dx_colsum,
                                 int drop_mode, float drop_p, const uint64_t* drop_rng, uint32_t drop_site, void* stream);
/* Keep bits of one attention layer: keep(b,h,q,k) = word (k & 3) of philox(seed; {k >> 2, (b H + h) S + q, site, step})
 * >= round(p 2^32), stored in the tile coordinates of the tcgen05 attention kernels (G = largest power of two <= 8 with
 * G S <= 128 dividing H heads share a tile; tile row r = q G + (h mod G); word [((b H/G + h/G) S G + r) ceil(S G / 32) + w]
 * bit j = tile column 32 w + j = key (32 w + j) / G of head (32 w + j) mod G).  simseg_attn_dropout_mask_words gives the
 * buffer size in 32-bit words (0: shape unsupported, S > 256). */
int64_t simseg_attn_dropout_mask_words(int B, int H, int S);
int simseg_attn_dropout_mask(simseg_ctx* ctx, int B, int H, int S, float drop_p, const uint64_t* drop_rng,
                             uint32_t drop_site, uint32_t* mask, int64_t mask_words, void* stream);
/* simseg_attention_fwd / _bwd with dropout(softmax(..)) (HF BertSelfAttention): out = dropout(P) V; lse is that of the
 * undropped softmax.  tcgen05 kernels only (S <= 224); other shapes return SIMSEG_ERR_UNSUPPORTED — there is no fallback. */
int simseg_attention_fwd_dropout(simseg_ctx* ctx, const void* q, const void* k, const void* v, int64_t stride_b,
                                 int64_t stride_s, int64_t stride_h, int B, int H, int S, const int32_t* key_len,
                                 float scale, void* out, float* lse, const uint32_t* drop_mask, float drop_p, void* stream);
int simseg_attention_bwd_dropout(simseg_ctx* ctx, const void* q, const void* k, const void* v, const void* out,
                                 const void* dout, const float* lse, int64_t stride_b, int64_t stride_s, int64_t stride_h,
                                 int B, int H, int S, const int32_t* key_len, float scale, void* dq, void* dk, void* dv,
                                 const uint32_t* drop_mask, float drop_p, void* stream);

/* ---- ViT / BERT embedding stages ------------------------------------------------------------ */
/* image [B,3,Hi,Wi] f32 NCHW -> patches bf16 [B*N, 768] (k = c*256 + py*16 + px), the im2col of the
 * stride-16 conv in timm PatchEmbed (vit_builder.py:14). */
int simseg_im2col16(simseg_ctx* ctx, const float* image, int B, int Hi, int Wi, void* patches, void* stream);
/* x[b,0,:] = cls + pos[0]; x[b,1+n,:] = patch[b*N+n,:] + pos[1+n]   (vit_builder.py:15-17) */
int simseg_vit_tokens_fwd(simseg_ctx* ctx, const void* patch, int patch_dtype, const float* cls,
                          const float* pos, int B, int N, int D, float* x, void* stream);
/* dpatch(bf16)[b*N+n] = dx[b,1+n]; dpos += sum_b dx[b]; dcls += sum_b dx[b,0]  (accumulating) */
int simseg_vit_tokens_bwd(simseg_ctx* ctx, const float* dx, int B, int N, int D, void* dpatch, float* dpos,
                          float* dcls, void* stream);
/* e[b,t,:] = word[ids[b,t]] + pos[t] + type[0]   (HF BertEmbeddings) */
int simseg_bert_embed_fwd(simseg_ctx* ctx, const int64_t* ids, const float* word, const float* pos,
                          const float* type0, int B, int T, int D, float* e, void* stream);
/* scatter-add of de into dword / dpos / dtype0 (accumulating, fp32 atomics) */
int simseg_bert_embed_bwd(simseg_ctx* ctx, const int64_t* ids, const float* de, int B, int T, int D,
                          float* dword, float* dpos, float* dtype0, void* stream);

/* ---- heads ----------------------------------------------------------------------------------- */
/* TopKPooling (components/pooling.py:42-65): per (sample, channel) mean of the top-k values over
 * tokens [tok_begin, tok_begin+ntok) of x[B,S,E] (row stride given); masked variant: tokens with
 * attention_mask==0 are replaced by -10000 (pooling.py:60).  k <= 8.  Saves the selected token
 * indices int32 [B,k,E] for backward.  Fused with L2norm (components/normalization.py:6-11) when
 * emb != NULL: emb = pooled / (||pooled|| + eps).  pooled f32 [B,E] always written. */
int simseg_topk_pool_l2norm_fwd(simseg_ctx* ctx, const void* x, int x_dtype, int B, int S, int E, int tok_begin,
                                int ntok, int k, const int64_t* attention_mask, int mask_ld, float eps,
                                float* pooled, float* emb, int32_t* sel_idx, void* stream);
/* backward of the fused head: demb [B,E] -> dx [B,S,E] (bf16, ZEROED then the k selected tokens of
 * every channel get dpooled/k). */
int simseg_topk_pool_l2norm_bwd(simseg_ctx* ctx, const float* demb, const float* pooled, const int32_t* sel_idx,
                                int B, int S, int E, int tok_begin, int k, float eps, int has_l2norm, void* dx,
                                void* stream);

/* Fused LoDA head: image_pool(image_projection(feat)) / text_pool(text_projection(feat), mask) of
 * pipelines/clip.py:87-93,111-120 = SimpleProjection (components/projection.py:45-46, y = x W^T, no bias) followed by
 * TopKPooling (components/pooling.py:57-65) [+ L2norm, components/normalization.py:6-11] as ONE tensor-core GEMM whose
 * epilogue keeps, per (sample, channel), the k largest bf16-rounded projections over tokens [tok_begin, tok_begin+ntok)
 * (earliest token first on ties; masked tokens = -10000): the [B,S,E] projection is never written.
 *   x  bf16 [B,S,D] (all S tokens of every sample, contiguous), w bf16 [E,D]; S <= 256, D % 64 == 0, E % 128 == 0, k <= 8
 *   pooled f32 [B,E]; emb f32 [B,E] or NULL (no L2norm); sel_idx int32 [B,k,E] or NULL (absolute token positions) */
int simseg_proj_topk_fwd(simseg_ctx* ctx, const void* x, const void* w, int B, int S, int D, int E, int tok_begin,
                         int ntok, int k, const int64_t* attention_mask, int mask_ld, float eps, float* pooled,
                         float* emb, int32_t* sel_idx, void* stream);
/* Backward of the fused head without the dense [B,S,E] gradient: demb [B,E] -> gy = dL/dpooled / k (workspace f32
 * [B,E], written), then   dx f32 [B,S,D] = dY W   (every row written; tokens nobody selected get zeros; NULL = skip;
 * needs wt = bf16 [D,E], the transposed weight copy)   and   dw f32 [E,D] += dY^T x   (NULL = skip; ACCUMULATES),
 * where dY[b,s,e] = gy[b,e] for the k selected tokens s of (b,e) and 0 elsewhere is generated tile by tile in shared
 * memory as the tcgen05 operand.  D % 128 == 0. */
int simseg_proj_topk_bwd(simseg_ctx* ctx, const float* demb, const float* pooled, const int32_t* sel_idx, const void* x,
                         const void* wt, int B, int S, int D, int E, int k, float eps, int has_l2norm, float* gy,
                         float* dx, float* dw, void* stream);

/* ---- InfoNCE (criteria/losses/mml_loss.py:51-96, driven twice by clip.py:123-149) ------------ */
#define SIMSEG_PREC_FP32 0 /* exact fp32 FFMA products                       */
#define SIMSEG_PREC_TF32 1 /* tcgen05 kind::tf32 products, fp32 accumulation */
/* feat1 [b,E] fp32 local rows, feat2g [Bg,E] fp32 gathered columns, temperature: device scalar
 * (clamped to [0.001,0.5] inside, mml_loss.py:56), target of row i = row_offset + i (:75).
 * cos_ws [b,Bg] fp32 (caller-owned): receives feat1 @ feat2g^T and must be handed unchanged to
 * simseg_infonce_bwd.  Outputs: loss_rows[b] = CE_i ; lse[b] ; argmax[b] (int32, first max, for the
 * top-1 accuracy of utils/misc.py:462-477, may be NULL) ; logits_out [b,Bg] optional. */
int simseg_infonce_fwd(simseg_ctx* ctx, const float* feat1, const float* feat2g, int b, int Bg, int E,
                       const float* temperature, int row_offset, int precision, float* cos_ws, float* logits_out,
                       float* loss_rows, float* lse, int32_t* argmax, void* stream);
/* grad_scale = dLoss/d(CE_i) (0.5/b for clip.py:140 with the mean of :89-91).  cos_ws is consumed
 * (overwritten by dLoss/dcos).  dfeat1 [b,E] written; dfeat2g [Bg,E] ACCUMULATED (GatherLayer.backward
 * sums over ranks, utils/dist.py:348-354); dtemp scalar ACCUMULATED (zero outside the clamp range). */
int simseg_infonce_bwd(simseg_ctx* ctx, const float* feat1, const float* feat2g, int b, int Bg, int E,
                       const float* temperature, int row_offset, int precision, const float* lse, float grad_scale,
                       float* cos_ws, float* dfeat1, float* dfeat2g, float* dtemp, void* stream);

/* The same loss on the tcgen05 tensor cores at fp32-grade accuracy, with the [b,Bg] score matrix consumed inside the
 * GEMM epilogue instead of being written (mml_loss.py:56,73-77 + utils/misc.py:462-477 in one pass).  fp32 operands are
 * split into bf16 hi + lo and multiplied as hi*hi + hi*lo + lo*hi (fp32 accumulate): cosine error <= ~1e-5, logits within
 * the 1e-3 bar at temperature 0.02.  `workspace` (simseg_infonce_fused_workspace_bytes, 256-byte aligned, caller-owned)
 * keeps the split operands and must be handed unchanged from _fwd to _bwd; _bwd recomputes the scores, emits
 * dLoss/dcos as split bf16 operands into the workspace and runs the two gradient GEMMs.  Outputs as above:
 * dfeat1 [b,E] written, dfeat2g [Bg,E] ACCUMULATED, dtemp ACCUMULATED (any of the three may be NULL). */
int64_t simseg_infonce_fused_workspace_bytes(int b, int Bg, int E);
int simseg_infonce_fused_fwd(simseg_ctx* ctx, const float* feat1, const float* feat2g, int b, int Bg, int E,
                             const float* temperature, int row_offset, void* workspace, int64_t workspace_bytes,
                             float* loss_rows, float* lse, int32_t* argmax, void* stream);
int simseg_infonce_fused_bwd(simseg_ctx* ctx, int b, int Bg, int E, const float* temperature, int row_offset,
                             const float* lse, float grad_scale, void* workspace, int64_t workspace_bytes, float* dfeat1,
                             float* dfeat2g, float* dtemp, void* stream);

/* ---- dense patch-text similarity map (tools/seg_evaluation.py:111-112,136-139) -------------- */
/* patches [rows,E] (rows = B*N projected patch tokens; f32 or bf16), text [C,E] same dtype
 * (class embeddings, already unit norm — seg_evaluation.py:71-72).
 * sim[rows,C] fp32 = (patch / max(||patch||,1e-12)) . text ; argmax[rows] int32 optional.
 * normalize = 0 skips the row normalisation (plain patches @ text^T). */
int64_t simseg_patch_text_sim_workspace_bytes(int64_t rows);
int simseg_patch_text_sim(simseg_ctx* ctx, const void* patches, int dtype, int64_t rows, int E, const void* text,
                          int C, int normalize, float* sim, int32_t* argmax, void* workspace, int64_t workspace_bytes,
                          void* stream);

/* ---- all-pairs retrieval similarity (tasks/clip/hooks/utils.py:36) --------------------------- */
/* out[M,Nr] fp32 = left[M,E] @ right[Nr,E]^T  (fp32 in, fp32 accumulate; precision = SIMSEG_PREC_*). */
int simseg_allpairs_sim(simseg_ctx* ctx, const float* left, const float* right, int M, int Nr, int E, int precision,
                        float* out, void* stream);
/* rank of the first right item whose gid matches the row's gid under a stable descending sort of
 * the similarities (tasks/clip/hooks/utils.py:37-42,63-65) without materialising the sort:
 * rank[i] = #{j : s_ij > s*_i} + #{j < j* : s_ij == s*_i}, (s*,j*) = best matching item. -1 if none. */
int simseg_retrieval_rank(simseg_ctx* ctx, const float* sim, int M, int Nr, const int64_t* left_gid,
                          const int64_t* right_gid, int32_t* rank, void* stream);

/* EmbANN._ann + RetrievalMetric's first-match rank (tasks/clip/hooks/utils.py:35-42,63-65) in one pass on the tensor
 * cores: the [M,Nr] similarity matrix (500 MB at 5k x 25k) and the [M,Nr] int64 argsort (1 GB) are never materialised.
 * Same split-bf16 products as simseg_infonce_fused_*; rank[i] as defined for simseg_retrieval_rank (-1: no right item
 * shares the row's group id). */
int64_t simseg_retrieval_fused_workspace_bytes(int M, int Nr, int E);
int simseg_retrieval_rank_fused(simseg_ctx* ctx, const float* left, const float* right, int M, int Nr, int E,
                                const int64_t* left_gid, const int64_t* right_gid, void* workspace, int64_t workspace_bytes,
                                int32_t* rank, void* stream);
/* out[M,Nr] fp32 = left @ right^T through the same split products (materialising variant of tasks/clip/hooks/utils.py:36;
 * workspace as simseg_retrieval_fused_workspace_bytes). */
int simseg_allpairs_sim_split(simseg_ctx* ctx, const float* left, const float* right, int M, int Nr, int E, void* workspace,
                              int64_t workspace_bytes, float* out, void* stream);

/* ---- zero-shot segmentation glue around the map (tools/seg_evaluation.py) ---- */
/* class embedding of the zero-shot classifier (seg_evaluation.py:71-72): out[c,:] = mean_p prompt[c,p,:] / ||mean||
 * (no eps).  prompt [C,P,E] fp32 = forward_text_project outputs of the P prompts of each class; E <= 1024. */
int simseg_seg_class_embed(simseg_ctx* ctx, const float* prompt, int C, int P, int E, float* out, void* stream);
/* image-level class selection (seg_evaluation.py:119-128,141-144): scores[b,c] = img[b,:] . text[c,:];
 * top-`topk` scores; threshold[b] = mean + std (unbiased) of those; cand[b, 0..max_cand) = the classes among the first
 * `max_cand` of the top-k, in order, skipping class ids 0 and 255, stopping at the first score < threshold; -1 pads.
 * scores / threshold may be NULL.  C <= 1024. */
int simseg_seg_select(simseg_ctx* ctx, const float* img_emb, const float* text_emb, int B, int C, int E, int topk,
                      int max_cand, float* scores, int32_t* cand, float* threshold, void* stream);
/* per (image, candidate): column cand[b,k] of sim [B,N,C] reshaped (h,w), nearest x`scale` up-sampling, min-max
 * normalisation (seg_evaluation.py:136-139,146-147) -> out [B,K,h*scale,w*scale] fp32 (zeros where cand == -1). */
int simseg_seg_upsample_norm(simseg_ctx* ctx, const float* sim, const int32_t* cand, int B, int N, int C, int K, int h,
                             int w, int scale, float* out, void* stream);
/* position-embedding resize at checkpoint load (utils/interpolate_pe.py:4-27, called from seg_evaluation.py:228-230):
 * src [num_extra + grid_src^2, D] fp32 -> dst [num_extra + grid_dst^2, D] fp32; the extra (class) tokens are copied, the
 * grid part is resized like F.interpolate(mode="bicubic", align_corners=False). */
int simseg_pos_embed_bicubic(simseg_ctx* ctx, const float* src, float* dst, int grid_src, int grid_dst, int D, int num_extra,
                             void* stream);

/* ---- debug (measurement only; synchronises the device) -------------------------------------------------------- */
/* Per-warp event trace of CTA 0 of the tcgen05 attention-backward kernel: 5 warps x 2048 events of (clock64 << 8 | id).
 * enable(1) clears and arms it, read() copies it to the host and returns the number of words written. */
int simseg_debug_trace_enable(int on);
int simseg_debug_trace_read(uint64_t* host, int n);

#ifdef __cplusplus
}
#endif
#endif /* SIMSEG_B200_H_ */
