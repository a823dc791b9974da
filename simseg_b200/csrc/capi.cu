// extern "C" surface of libsimseg_b200 (see include/simseg_b200.h).  Thin argument checks + dispatch.
#include "common.cuh"

#include <cstring>
#include <cstdlib>

namespace simseg {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// implemented in the kernel translation units
int gemm_impl(Ctx*, const simseg_gemm_args*, cudaStream_t);
int cast_bf16_impl(Ctx*, const float*, void*, void*, int64_t, int64_t, cudaStream_t);
int cast_bf16_multi_impl(Ctx*, const simseg_cast_item*, int, int64_t, cudaStream_t);
int colsum_impl(Ctx*, const void*, int, int64_t, int64_t, int64_t, float*, int, cudaStream_t);
int gelu_fwd_impl(Ctx*, const void*, void*, int64_t, cudaStream_t);
int layernorm_fwd_impl(Ctx*, const void*, int, const float*, const float*, float, int64_t, int, void*, float*, float*, float*, const void*, float*, float, const void*, uint32_t, cudaStream_t);
int layernorm_bwd_impl(Ctx*, const void*, int, const float*, const void*, int, const float*, const float*, const float*, int64_t, int, float*, int, void*, float*, float*, float*, int, float, const void*, uint32_t, cudaStream_t);
int attention_fwd_impl(Ctx*, const void*, const void*, const void*, int64_t, int64_t, int64_t, int, int, int, const int32_t*, float, void*, float*, cudaStream_t);
int attention_bwd_impl(Ctx*, const void*, const void*, const void*, const void*, const void*, const float*, int64_t, int64_t, int64_t, int, int, int, const int32_t*, float, void*, void*, void*, cudaStream_t);
int im2col16_impl(Ctx*, const float*, int, int, int, void*, cudaStream_t);
int vit_tokens_fwd_impl(Ctx*, const void*, int, const float*, const float*, int, int, int, float*, cudaStream_t);
int vit_tokens_bwd_impl(Ctx*, const float*, int, int, int, void*, float*, float*, cudaStream_t);
int bert_embed_fwd_impl(Ctx*, const int64_t*, const float*, const float*, const float*, int, int, int, float*, cudaStream_t);
int bert_embed_bwd_impl(Ctx*, const int64_t*, const float*, int, int, int, float*, float*, float*, cudaStream_t);
int topk_pool_l2norm_fwd_impl(Ctx*, const void*, int, int, int, int, int, int, int, const int64_t*, int, float, float*, float*, int32_t*, cudaStream_t);
int topk_pool_l2norm_bwd_impl(Ctx*, const float*, const float*, const int32_t*, int, int, int, int, float, int, void*, cudaStream_t);
int proj_topk_fwd_impl(Ctx*, const void*, const void*, int, int, int, int, int, int, int, const int64_t*, int, float, float*, float*, int32_t*, cudaStream_t);
int proj_topk_bwd_impl(Ctx*, const float*, const float*, const int32_t*, const void*, const void*, int, int, int, int, int, float, int, float*, float*, float*, cudaStream_t);
int sgemm_impl(Ctx*, const float*, const float*, float*, int, int, int, int64_t, int64_t, int64_t, int, int, int, cudaStream_t);
int nce_rows_fwd_impl(Ctx*, const float*, int, int, int64_t, const float*, int, float*, float*, float*, int32_t*, cudaStream_t);
int nce_rows_bwd_impl(Ctx*, float*, int, int, int64_t, const float*, int, const float*, float, float*, cudaStream_t);
int row_inv_norm_impl(Ctx*, const void*, int, int64_t, int, float*, cudaStream_t);
int row_argmax_impl(Ctx*, const float*, int64_t, int, int32_t*, cudaStream_t);
int retrieval_rank_impl(Ctx*, const float*, int, int, const int64_t*, const int64_t*, int32_t*, cudaStream_t);
int attention_bwd_tc_impl(Ctx*, const void*, const void*, const void*, const void*, const void*, const float*, int64_t, int64_t, int64_t, int, int, int, const int32_t*, float, void*, void*, void*, const uint32_t*, float, cudaStream_t);
int attention_fwd_tc_impl(Ctx*, const void*, const void*, const void*, int64_t, int64_t, int64_t, int, int, int, const int32_t*, float, void*, float*, const uint32_t*, float, cudaStream_t);
int64_t attn_dropout_mask_words_impl(int, int, int);
int attn_dropout_mask_impl(Ctx*, int, int, int, float, const void*, uint32_t, uint32_t*, int64_t, cudaStream_t);
int seg_class_embed_impl(Ctx*, const float*, int, int, int, float*, cudaStream_t);
int seg_select_impl(Ctx*, const float*, const float*, int, int, int, int, int, float*, int32_t*, float*, cudaStream_t);
int seg_upsample_norm_impl(Ctx*, const float*, const int32_t*, int, int, int, int, int, int, int, float*, cudaStream_t);
int pos_embed_bicubic_impl(Ctx*, const float*, float*, int, int, int, int, cudaStream_t);
int patch_sim_fused_impl(Ctx*, const void*, int64_t, int, const void*, int, int, float*, int32_t*, cudaStream_t);
int debug_trace_enable_impl(int);
int debug_trace_read_impl(unsigned long long*, int);
int64_t infonce_fused_workspace_bytes(int, int, int);
int infonce_fused_fwd_impl(Ctx*, const float*, const float*, int, int, int, const float*, int, void*, int64_t, float*, float*, int32_t*, cudaStream_t);
int infonce_fused_bwd_impl(Ctx*, int, int, int, const float*, int, const float*, float, void*, int64_t, float*, float*, float*, cudaStream_t);
int64_t retrieval_fused_workspace_bytes(int, int, int);
int retrieval_rank_fused_impl(Ctx*, const float*, const float*, int, int, int, const int64_t*, const int64_t*, void*, int64_t, int32_t*, cudaStream_t);
int allpairs_split_impl(Ctx*, const float*, const float*, int, int, int, void*, int64_t, float*, cudaStream_t);

// fp32 product in the requested precision: C[M,N] (+)= A(m,k) B(n,k)
static int fp32_matmul(Ctx* c, const float* A, const float* B, float* C, int M, int N, int K, int64_t lda, int64_t ldb,
                       int64_t ldc, int a_major, int b_major, int accumulate, int precision, cudaStream_t st) {
  // tensor-core path: K-major tf32 operands with 16-byte aligned rows; everything else runs exact fp32 FFMA
  const bool tc_ok = precision == SIMSEG_PREC_TF32 && !a_major && !b_major && lda % 4 == 0 && ldb % 4 == 0 &&
                     (reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0;
  if (!tc_ok) return sgemm_impl(c, A, B, C, M, N, K, lda, ldb, ldc, a_major, b_major, accumulate, st);
  simseg_gemm_args g;
  memset(&g, 0, sizeof(g));
  g.a = A; g.b = B; g.d = C; g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldb = ldb; g.ldd = ldc;
  g.a_major = a_major; g.b_major = b_major; g.in_dtype = SIMSEG_F32; g.out_dtype = SIMSEG_F32;
  g.epilogue = SIMSEG_EPI_NONE; g.accumulate = accumulate;
  return gemm_impl(c, &g, st);
}

}  // namespace simseg

using namespace simseg;

struct simseg_ctx { Ctx c; };
#define CTX_OR_FAIL()                                   \
  if (ctx == nullptr) { set_error("null ctx"); return SIMSEG_ERR_INVALID; } \
  Ctx* c = &ctx->c;                                     \
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream)

extern "C" {

int simseg_version(void) { return 100; }
const char* simseg_last_error(void) { return g_err; }

int simseg_ctx_create(int device, simseg_ctx** out) {
  if (!out) { set_error("null out"); return SIMSEG_ERR_INVALID; }
  cudaDeviceProp prop;
  SIMSEG_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; libsimseg_b200 carries sm_100a code only", device, prop.major, prop.minor);
    return SIMSEG_ERR_UNSUPPORTED;
  }
  SIMSEG_CUDA(cudaSetDevice(device));
  simseg_ctx* x = new simseg_ctx();
  x->c.device = device;
  x->c.num_sms = prop.multiProcessorCount;
  x->c.launches = 0;
  *out = x;
  return SIMSEG_OK;
}
int simseg_ctx_destroy(simseg_ctx* ctx) { delete ctx; return SIMSEG_OK; }
int64_t simseg_ctx_launch_count(simseg_ctx* ctx, int reset) {
  if (!ctx) return -1;
  const int64_t n = ctx->c.launches;
  if (reset) ctx->c.launches = 0;
  return n;
}

int simseg_gemm(simseg_ctx* ctx, const simseg_gemm_args* args, void* stream) {
  CTX_OR_FAIL();
  if (!args) { set_error("null args"); return SIMSEG_ERR_INVALID; }
  return gemm_impl(c, args, st);
}
int simseg_cast_bf16_multi(simseg_ctx* ctx, const simseg_cast_item* items, int n_items, int64_t total_blocks, void* stream) {
  CTX_OR_FAIL();
  return cast_bf16_multi_impl(c, items, n_items, total_blocks, st);
}
int simseg_cast_bf16(simseg_ctx* ctx, const float* src, void* dst, void* dst_t, int64_t rows, int64_t cols, void* stream) {
  CTX_OR_FAIL();
  return cast_bf16_impl(c, src, dst, dst_t, rows, cols, st);
}
int simseg_colsum(simseg_ctx* ctx, const void* x, int dtype, int64_t M, int64_t N, int64_t ldx, float* out, int accumulate, void* stream) {
  CTX_OR_FAIL();
  return colsum_impl(c, x, dtype, M, N, ldx, out, accumulate, st);
}
int simseg_gelu_fwd(simseg_ctx* ctx, const void* h, void* a, int64_t n, void* stream) {
  CTX_OR_FAIL();
  return gelu_fwd_impl(c, h, a, n, st);
}
int simseg_layernorm_fwd(simseg_ctx* ctx, const void* x, int x_dtype, const float* gamma, const float* beta, float eps,
                         int64_t M, int D, void* y_bf16, float* y_f32, float* mean, float* rstd, void* stream) {
  CTX_OR_FAIL();
  return layernorm_fwd_impl(c, x, x_dtype, gamma, beta, eps, M, D, y_bf16, y_f32, mean, rstd, nullptr, nullptr, 0.f, nullptr, 0u, st);
}
int simseg_add_layernorm_fwd(simseg_ctx* ctx, const float* x, const void* add_bf16, const float* gamma, const float* beta,
                             float eps, int64_t M, int D, float* sum_out, void* y_bf16, float* y_f32, float* mean, float* rstd,
                             void* stream) {
  CTX_OR_FAIL();
  SIMSEG_CHECK_ARG(x != nullptr && add_bf16 != nullptr, "add_layernorm_fwd: x and add are required");
  return layernorm_fwd_impl(c, x, SIMSEG_F32, gamma, beta, eps, M, D, y_bf16, y_f32, mean, rstd, add_bf16, sum_out, 0.f, nullptr, 0u, st);
}
int simseg_layernorm_fwd_dropout(simseg_ctx* ctx, const float* x, const void* add_bf16, const float* gamma, const float* beta,
                                 float eps, int64_t M, int D, float* sum_out, void* y_bf16, float* y_f32, float* mean,
                                 float* rstd, float drop_p, const uint64_t* drop_rng, uint32_t drop_site, void* stream) {
  CTX_OR_FAIL();
  SIMSEG_CHECK_ARG(x != nullptr, "layernorm_fwd_dropout: x is required");
  SIMSEG_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "layernorm_fwd_dropout: p=%g outside [0,1)", drop_p);
  return layernorm_fwd_impl(c, x, SIMSEG_F32, gamma, beta, eps, M, D, y_bf16, y_f32, mean, rstd, add_bf16, sum_out, drop_p,
                            drop_rng, drop_site, st);
}
int simseg_layernorm_bwd(simseg_ctx* ctx, const void* dy, int dy_dtype, const float* dy2, const void* x, int x_dtype,
                         const float* gamma, const float* mean, const float* rstd, int64_t M, int D, float* dx,
                         int dx_accumulate, void* dx_bf16, float* dgamma, float* dbeta, float* dx_colsum, void* stream) {
  CTX_OR_FAIL();
  return layernorm_bwd_impl(c, dy, dy_dtype, dy2, x, x_dtype, gamma, mean, rstd, M, D, dx, dx_accumulate, dx_bf16, dgamma,
                            dbeta, dx_colsum, 0, 0.f, nullptr, 0u, st);
}
int simseg_layernorm_bwd_dropout(simseg_ctx* ctx, const void* dy, int dy_dtype, const float* dy2, const float* x,
                                 const float* gamma, const float* mean, const float* rstd, int64_t M, int D, float* dx,
                                 int dx_accumulate, void* dx_bf16, float* dgamma, float* dbeta, float* dx_colsum, int drop_mode,
                                 float drop_p, const uint64_t* drop_rng, uint32_t drop_site, void* stream) {
  CTX_OR_FAIL();
  SIMSEG_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "layernorm_bwd_dropout: p=%g outside [0,1)", drop_p);
  return layernorm_bwd_impl(c, dy, dy_dtype, dy2, x, SIMSEG_F32, gamma, mean, rstd, M, D, dx, dx_accumulate, dx_bf16, dgamma,
                            dbeta, dx_colsum, drop_mode, drop_p, drop_rng, drop_site, st);
}
int simseg_attention_fwd(simseg_ctx* ctx, const void* q, const void* k, const void* v, int64_t stride_b, int64_t stride_s,
                         int64_t stride_h, int B, int H, int S, const int32_t* key_len, float scale, void* out, float* lse,
                         void* stream) {
  CTX_OR_FAIL();
  // tcgen05 kernel for S <= 224 (attention_sm100.cu); shapes it rejects run the mma.sync kernel.
  // SIMSEG_ATTN_FWD=mma|tc|ts overrides the choice (tests exercise all three; default ts = P kept in tensor memory).
  const char* force = getenv("SIMSEG_ATTN_FWD");
  const bool want_tc = force ? (force[0] == 't') : true;       // short sequences are packed several heads per tile
  if (want_tc) {
    const int rc = attention_fwd_tc_impl(c, q, k, v, stride_b, stride_s, stride_h, B, H, S, key_len, scale, out, lse, nullptr, 0.f, st);
    if (rc != SIMSEG_ERR_UNSUPPORTED) return rc;
  }
  return attention_fwd_impl(c, q, k, v, stride_b, stride_s, stride_h, B, H, S, key_len, scale, out, lse, st);
}
int simseg_attention_bwd(simseg_ctx* ctx, const void* q, const void* k, const void* v, const void* out, const void* dout,
                         const float* lse, int64_t stride_b, int64_t stride_s, int64_t stride_h, int B, int H, int S,
                         const int32_t* key_len, float scale, void* dq, void* dk, void* dv, void* stream) {
  CTX_OR_FAIL();
  // tcgen05 kernel for S <= 256 (attention_sm100.cu); longer sequences and shapes it rejects run the mma.sync kernel.
  // SIMSEG_ATTN_BWD=mma|tc overrides the choice (tests exercise both).
  const char* force = getenv("SIMSEG_ATTN_BWD");
  const bool want_tc = force ? (force[0] == 't') : true;       // short sequences are packed several heads per tile
  if (want_tc) {
    const int rc = attention_bwd_tc_impl(c, q, k, v, out, dout, lse, stride_b, stride_s, stride_h, B, H, S, key_len, scale, dq, dk, dv, nullptr, 0.f, st);
    if (rc != SIMSEG_ERR_UNSUPPORTED) return rc;
  }
  return attention_bwd_impl(c, q, k, v, out, dout, lse, stride_b, stride_s, stride_h, B, H, S, key_len, scale, dq, dk, dv, st);
}
int64_t simseg_attn_dropout_mask_words(int B, int H, int S) { return attn_dropout_mask_words_impl(B, H, S); }
int simseg_attn_dropout_mask(simseg_ctx* ctx, int B, int H, int S, float drop_p, const uint64_t* drop_rng, uint32_t drop_site,
                             uint32_t* mask, int64_t mask_words, void* stream) {
  CTX_OR_FAIL();
  SIMSEG_CHECK_ARG(drop_p > 0.f && drop_p < 1.f && drop_rng != nullptr && mask != nullptr,
                   "attn_dropout_mask: needs 0 < p < 1, a device {seed, step} pair and a mask buffer");
  return attn_dropout_mask_impl(c, B, H, S, drop_p, drop_rng, drop_site, mask, mask_words, st);
}
// Attention with dropout on the probabilities runs on the tcgen05 kernels only (S <= 224 forward, <= 256 backward, 16-byte
// aligned head rows): there is no fallback, shapes they reject are an error.
int simseg_attention_fwd_dropout(simseg_ctx* ctx, const void* q, const void* k, const void* v, int64_t stride_b,
                                 int64_t stride_s, int64_t stride_h, int B, int H, int S, const int32_t* key_len, float scale,
                                 void* out, float* lse, const uint32_t* drop_mask, float drop_p, void* stream) {
  CTX_OR_FAIL();
  SIMSEG_CHECK_ARG(drop_mask != nullptr && drop_p > 0.f && drop_p < 1.f, "attention_fwd_dropout: needs a mask and 0 < p < 1");
  const int rc = attention_fwd_tc_impl(c, q, k, v, stride_b, stride_s, stride_h, B, H, S, key_len, scale, out, lse, drop_mask, drop_p, st);
  if (rc == SIMSEG_ERR_UNSUPPORTED) set_error("attention_fwd_dropout: shape B=%d H=%d S=%d not supported with dropout (S <= 224, aligned strides)", B, H, S);
  return rc;
}
int simseg_attention_bwd_dropout(simseg_ctx* ctx, const void* q, const void* k, const void* v, const void* out,
                                 const void* dout, const float* lse, int64_t stride_b, int64_t stride_s, int64_t stride_h, int B,
                                 int H, int S, const int32_t* key_len, float scale, void* dq, void* dk, void* dv,
                                 const uint32_t* drop_mask, float drop_p, void* stream) {
  CTX_OR_FAIL();
  SIMSEG_CHECK_ARG(drop_mask != nullptr && drop_p > 0.f && drop_p < 1.f, "attention_bwd_dropout: needs a mask and 0 < p < 1");
  const int rc = attention_bwd_tc_impl(c, q, k, v, out, dout, lse, stride_b, stride_s, stride_h, B, H, S, key_len, scale, dq, dk, dv, drop_mask, drop_p, st);
  if (rc == SIMSEG_ERR_UNSUPPORTED) set_error("attention_bwd_dropout: shape B=%d H=%d S=%d not supported with dropout (S <= 224, aligned strides)", B, H, S);
  return rc;
}
int simseg_im2col16(simseg_ctx* ctx, const float* image, int B, int Hi, int Wi, void* patches, void* stream) {
  CTX_OR_FAIL();
  return im2col16_impl(c, image, B, Hi, Wi, patches, st);
}
int simseg_vit_tokens_fwd(simseg_ctx* ctx, const void* patch, int patch_dtype, const float* cls, const float* pos, int B,
                          int N, int D, float* x, void* stream) {
  CTX_OR_FAIL();
  return vit_tokens_fwd_impl(c, patch, patch_dtype, cls, pos, B, N, D, x, st);
}
int simseg_vit_tokens_bwd(simseg_ctx* ctx, const float* dx, int B, int N, int D, void* dpatch, float* dpos, float* dcls,
                          void* stream) {
  CTX_OR_FAIL();
  return vit_tokens_bwd_impl(c, dx, B, N, D, dpatch, dpos, dcls, st);
}
int simseg_bert_embed_fwd(simseg_ctx* ctx, const int64_t* ids, const float* word, const float* pos, const float* type0,
                          int B, int T, int D, float* e, void* stream) {
  CTX_OR_FAIL();
  return bert_embed_fwd_impl(c, ids, word, pos, type0, B, T, D, e, st);
}
int simseg_bert_embed_bwd(simseg_ctx* ctx, const int64_t* ids, const float* de, int B, int T, int D, float* dword,
                          float* dpos, float* dtype0, void* stream) {
  CTX_OR_FAIL();
  return bert_embed_bwd_impl(c, ids, de, B, T, D, dword, dpos, dtype0, st);
}
int simseg_topk_pool_l2norm_fwd(simseg_ctx* ctx, const void* x, int x_dtype, int B, int S, int E, int tok_begin, int ntok,
                                int k, const int64_t* attention_mask, int mask_ld, float eps, float* pooled, float* emb,
                                int32_t* sel_idx, void* stream) {
  CTX_OR_FAIL();
  return topk_pool_l2norm_fwd_impl(c, x, x_dtype, B, S, E, tok_begin, ntok, k, attention_mask, mask_ld, eps, pooled, emb,
                                   sel_idx, st);
}
int simseg_topk_pool_l2norm_bwd(simseg_ctx* ctx, const float* demb, const float* pooled, const int32_t* sel_idx, int B,
                                int S, int E, int tok_begin, int k, float eps, int has_l2norm, void* dx, void* stream) {
  CTX_OR_FAIL();
  (void)tok_begin;   // sel_idx already holds absolute token positions
  return topk_pool_l2norm_bwd_impl(c, demb, pooled, sel_idx, B, S, E, k, eps, has_l2norm, dx, st);
}

int simseg_proj_topk_fwd(simseg_ctx* ctx, const void* x, const void* w, int B, int S, int D, int E, int tok_begin, int ntok,
                         int k, const int64_t* attention_mask, int mask_ld, float eps, float* pooled, float* emb,
                         int32_t* sel_idx, void* stream) {
  CTX_OR_FAIL();
  return proj_topk_fwd_impl(c, x, w, B, S, D, E, tok_begin, ntok, k, attention_mask, mask_ld, eps, pooled, emb, sel_idx, st);
}
int simseg_proj_topk_bwd(simseg_ctx* ctx, const float* demb, const float* pooled, const int32_t* sel_idx, const void* x,
                         const void* wt, int B, int S, int D, int E, int k, float eps, int has_l2norm, float* gy, float* dx,
                         float* dw, void* stream) {
  CTX_OR_FAIL();
  return proj_topk_bwd_impl(c, demb, pooled, sel_idx, x, wt, B, S, D, E, k, eps, has_l2norm, gy, dx, dw, st);
}

int simseg_infonce_fwd(simseg_ctx* ctx, const float* feat1, const float* feat2g, int b, int Bg, int E,
                       const float* temperature, int row_offset, int precision, float* cos_ws, float* logits_out,
                       float* loss_rows, float* lse, int32_t* argmax, void* stream) {
  CTX_OR_FAIL();
  SIMSEG_CHECK_ARG(cos_ws != nullptr, "infonce_fwd: cos_ws workspace required");
  int rc = fp32_matmul(c, feat1, feat2g, cos_ws, b, Bg, E, E, E, Bg, 0, 0, 0, precision, st);
  if (rc) return rc;
  return nce_rows_fwd_impl(c, cos_ws, b, Bg, Bg, temperature, row_offset, logits_out, loss_rows, lse, argmax, st);
}
int simseg_infonce_bwd(simseg_ctx* ctx, const float* feat1, const float* feat2g, int b, int Bg, int E,
                       const float* temperature, int row_offset, int precision, const float* lse, float grad_scale,
                       float* cos_ws, float* dfeat1, float* dfeat2g, float* dtemp, void* stream) {
  CTX_OR_FAIL();
  SIMSEG_CHECK_ARG(cos_ws != nullptr, "infonce_bwd: cos_ws workspace required");
  int rc = nce_rows_bwd_impl(c, cos_ws, b, Bg, Bg, temperature, row_offset, lse, grad_scale, dtemp, st);
  if (rc) return rc;
  // dfeat1[b,E] = G[b,Bg] @ feat2g[Bg,E]          (A K-major, B stored [K,N])
  if (dfeat1) {
    rc = fp32_matmul(c, cos_ws, feat2g, dfeat1, b, E, Bg, Bg, E, E, 0, 1, 0, precision, st);
    if (rc) return rc;
  }
  // dfeat2g[Bg,E] += G^T[Bg,b] @ feat1[b,E]       (A stored [K,M], B stored [K,N])
  if (dfeat2g) {
    rc = fp32_matmul(c, cos_ws, feat1, dfeat2g, Bg, E, b, Bg, E, E, 1, 1, 1, precision, st);
    if (rc) return rc;
  }
  return SIMSEG_OK;
}

int simseg_debug_trace_enable(int on) { return debug_trace_enable_impl(on); }
int simseg_debug_trace_read(uint64_t* host, int n) { return debug_trace_read_impl(reinterpret_cast<unsigned long long*>(host), n); }

int64_t simseg_infonce_fused_workspace_bytes(int b, int Bg, int E) { return infonce_fused_workspace_bytes(b, Bg, E); }
int simseg_infonce_fused_fwd(simseg_ctx* ctx, const float* feat1, const float* feat2g, int b, int Bg, int E,
                             const float* temperature, int row_offset, void* workspace, int64_t workspace_bytes,
                             float* loss_rows, float* lse, int32_t* argmax, void* stream) {
  CTX_OR_FAIL();
  SIMSEG_CHECK_ARG(feat1 && feat2g && temperature && loss_rows && lse, "infonce_fused_fwd: null pointer");
  return infonce_fused_fwd_impl(c, feat1, feat2g, b, Bg, E, temperature, row_offset, workspace, workspace_bytes, loss_rows, lse,
                                argmax, st);
}
int simseg_infonce_fused_bwd(simseg_ctx* ctx, int b, int Bg, int E, const float* temperature, int row_offset,
                             const float* lse, float grad_scale, void* workspace, int64_t workspace_bytes, float* dfeat1,
                             float* dfeat2g, float* dtemp, void* stream) {
  CTX_OR_FAIL();
  SIMSEG_CHECK_ARG(temperature && lse, "infonce_fused_bwd: null pointer");
  return infonce_fused_bwd_impl(c, b, Bg, E, temperature, row_offset, lse, grad_scale, workspace, workspace_bytes, dfeat1, dfeat2g,
                                dtemp, st);
}
int64_t simseg_retrieval_fused_workspace_bytes(int M, int Nr, int E) { return retrieval_fused_workspace_bytes(M, Nr, E); }
int simseg_retrieval_rank_fused(simseg_ctx* ctx, const float* left, const float* right, int M, int Nr, int E,
                                const int64_t* left_gid, const int64_t* right_gid, void* workspace, int64_t workspace_bytes,
                                int32_t* rank, void* stream) {
  CTX_OR_FAIL();
  SIMSEG_CHECK_ARG(left && right && left_gid && right_gid && rank, "retrieval_rank_fused: null pointer");
  return retrieval_rank_fused_impl(c, left, right, M, Nr, E, left_gid, right_gid, workspace, workspace_bytes, rank, st);
}
int simseg_allpairs_sim_split(simseg_ctx* ctx, const float* left, const float* right, int M, int Nr, int E, void* workspace,
                              int64_t workspace_bytes, float* out, void* stream) {
  CTX_OR_FAIL();
  SIMSEG_CHECK_ARG(left && right && out, "allpairs_sim_split: null pointer");
  return allpairs_split_impl(c, left, right, M, Nr, E, workspace, workspace_bytes, out, st);
}

int64_t simseg_patch_text_sim_workspace_bytes(int64_t rows) { return ((rows * 4 + 255) / 256) * 256; }

int simseg_patch_text_sim(simseg_ctx* ctx, const void* patches, int dtype, int64_t rows, int E, const void* text, int C,
                          int normalize, float* sim, int32_t* argmax, void* workspace, int64_t workspace_bytes,
                          void* stream) {
  CTX_OR_FAIL();
  SIMSEG_CHECK_ARG(rows > 0 && C > 0 && E > 0, "patch_text_sim: empty");
  if (dtype == SIMSEG_BF16) {
    // single-pass fused kernel (patch_sim.cu); shapes it does not cover take the generic 3-kernel sequence below
    const int frc = patch_sim_fused_impl(c, patches, rows, E, text, C, normalize, sim, argmax, st);
    if (frc != SIMSEG_ERR_UNSUPPORTED) return frc;
  }
  simseg_gemm_args g;
  memset(&g, 0, sizeof(g));
  g.a = patches; g.b = text; g.d = sim; g.M = rows; g.N = C; g.K = E; g.lda = E; g.ldb = E; g.ldd = C;
  g.in_dtype = dtype; g.out_dtype = SIMSEG_F32; g.epilogue = SIMSEG_EPI_NONE;
  int rc;
  if (normalize) {
    SIMSEG_CHECK_ARG(workspace != nullptr && workspace_bytes >= rows * 4, "patch_text_sim: workspace too small");
    rc = row_inv_norm_impl(c, patches, dtype, rows, E, reinterpret_cast<float*>(workspace), st);
    if (rc) return rc;
    g.epilogue = SIMSEG_EPI_ROWSCALE;
    g.row_scale = reinterpret_cast<const float*>(workspace);
  }
  rc = gemm_impl(c, &g, st);
  if (rc) return rc;
  if (argmax) return row_argmax_impl(c, sim, rows, C, argmax, st);
  return SIMSEG_OK;
}

int simseg_allpairs_sim(simseg_ctx* ctx, const float* left, const float* right, int M, int Nr, int E, int precision,
                        float* out, void* stream) {
  CTX_OR_FAIL();
  return fp32_matmul(c, left, right, out, M, Nr, E, E, E, Nr, 0, 0, 0, precision, st);
}
int simseg_retrieval_rank(simseg_ctx* ctx, const float* sim, int M, int Nr, const int64_t* left_gid,
                          const int64_t* right_gid, int32_t* rank, void* stream) {
  CTX_OR_FAIL();
  return retrieval_rank_impl(c, sim, M, Nr, left_gid, right_gid, rank, st);
}

int simseg_seg_class_embed(simseg_ctx* ctx, const float* prompt, int C, int P, int E, float* out, void* stream) {
  CTX_OR_FAIL();
  return seg_class_embed_impl(c, prompt, C, P, E, out, st);
}
int simseg_seg_select(simseg_ctx* ctx, const float* img_emb, const float* text_emb, int B, int C, int E, int topk, int max_cand,
                      float* scores, int32_t* cand, float* threshold, void* stream) {
  CTX_OR_FAIL();
  return seg_select_impl(c, img_emb, text_emb, B, C, E, topk, max_cand, scores, cand, threshold, st);
}
int simseg_seg_upsample_norm(simseg_ctx* ctx, const float* sim, const int32_t* cand, int B, int N, int C, int K, int h, int w,
                             int scale, float* out, void* stream) {
  CTX_OR_FAIL();
  return seg_upsample_norm_impl(c, sim, cand, B, N, C, K, h, w, scale, out, st);
}
int simseg_pos_embed_bicubic(simseg_ctx* ctx, const float* src, float* dst, int grid_src, int grid_dst, int D, int num_extra,
                             void* stream) {
  CTX_OR_FAIL();
  SIMSEG_CHECK_ARG(src && dst, "pos_embed_bicubic: null pointer");
  return pos_embed_bicubic_impl(c, src, dst, grid_src, grid_dst, D, num_extra, st);
}

}  // extern "C"
