// Embedding stages (ViT patchify / token assembly, BERT embeddings) and the LoDA head
// (top-k mean pooling fused with L2norm), forward and backward.
#include "common.cuh"

namespace simseg {

// ------------------------------------------------------------------------------------------------
// im2col for the stride-16 patch conv: image [B,3,Hi,Wi] f32 -> patches bf16 [B*N, 768], k = c*256 + py*16 + px.
// One thread converts 8 consecutive pixels (32 B read, 16 B write); reads are fully coalesced.
__global__ void im2col16_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int Hi, int Wi) {
  const int64_t chunks_per_row = Wi / 8;
  const int64_t total = static_cast<int64_t>(B) * 3 * Hi * chunks_per_row;
  const int pw = Wi / 16, ph = Hi / 16;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t xc = i % chunks_per_row;
    int64_t t = i / chunks_per_row;
    const int y = static_cast<int>(t % Hi); t /= Hi;
    const int c = static_cast<int>(t % 3);
    const int64_t b = t / 3;
    const float4 f0 = *reinterpret_cast<const float4*>(img + i * 8);
    const float4 f1 = *reinterpret_cast<const float4*>(img + i * 8 + 4);
    const int x = static_cast<int>(xc) * 8;
    const int64_t n = static_cast<int64_t>(y / 16) * pw + x / 16;
    const int k = c * 256 + (y % 16) * 16 + (x % 16);
    uint4 u;
    u.x = pack_bf16(f0.x, f0.y); u.y = pack_bf16(f0.z, f0.w); u.z = pack_bf16(f1.x, f1.y); u.w = pack_bf16(f1.z, f1.w);
    *reinterpret_cast<uint4*>(out + (b * ph * pw + n) * 768 + k) = u;
  }
}

int im2col16_impl(Ctx* ctx, const float* image, int B, int Hi, int Wi, void* patches, cudaStream_t st) {
  SIMSEG_CHECK_ARG(B > 0 && Hi % 16 == 0 && Wi % 16 == 0, "im2col16: image size must be a multiple of 16");
  const int64_t total = static_cast<int64_t>(B) * 3 * Hi * (Wi / 8);
  const int grid = static_cast<int>(imin64(cdiv(total, 256), static_cast<int64_t>(ctx->num_sms) * 16));
  im2col16_kernel<<<grid, 256, 0, st>>>(image, reinterpret_cast<__nv_bfloat16*>(patches), B, Hi, Wi);
  ctx->launches++;
  SIMSEG_LAUNCH_CHECK();
  return SIMSEG_OK;
}

// ------------------------------------------------------------------------------------------------
template <bool PBF16>
__global__ void vit_tokens_fwd_kernel(const void* __restrict__ patch, const float* __restrict__ cls,
                                      const float* __restrict__ pos, int B, int N, int D, float* __restrict__ x) {
  const int D4 = D / 4;
  const int64_t total = static_cast<int64_t>(B) * (N + 1) * D4;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int d = static_cast<int>(i % D4) * 4;
    const int64_t t = i / D4;
    const int s = static_cast<int>(t % (N + 1));
    const int64_t b = t / (N + 1);
    const float4 p = *reinterpret_cast<const float4*>(pos + static_cast<int64_t>(s) * D + d);
    float4 v;
    if (s == 0) {
      v = *reinterpret_cast<const float4*>(cls + d);
    } else if (PBF16) {
      const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(patch) + (b * N + s - 1) * D + d);
      v = make_float4(bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y));
    } else {
      v = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(patch) + (b * N + s - 1) * D + d);
    }
    *reinterpret_cast<float4*>(x + t * D + d) = make_float4(v.x + p.x, v.y + p.y, v.z + p.z, v.w + p.w);
  }
}

int vit_tokens_fwd_impl(Ctx* ctx, const void* patch, int patch_dtype, const float* cls, const float* pos, int B, int N,
                        int D, float* x, cudaStream_t st) {
  SIMSEG_CHECK_ARG(B > 0 && N > 0 && D % 4 == 0, "vit_tokens_fwd: bad shape");
  const int64_t total = static_cast<int64_t>(B) * (N + 1) * (D / 4);
  const int grid = static_cast<int>(imin64(cdiv(total, 256), static_cast<int64_t>(ctx->num_sms) * 16));
  if (patch_dtype == SIMSEG_BF16) vit_tokens_fwd_kernel<true><<<grid, 256, 0, st>>>(patch, cls, pos, B, N, D, x);
  else vit_tokens_fwd_kernel<false><<<grid, 256, 0, st>>>(patch, cls, pos, B, N, D, x);
  ctx->launches++;
  SIMSEG_LAUNCH_CHECK();
  return SIMSEG_OK;
}

// sum over the batch: out[s, d] += sum_b dx[b, s, d]; also emits the bf16 patch gradient.
// grid = (S, batch_splits); block = D/4 threads (float4 lanes).
__global__ void vit_tokens_bwd_kernel(const float* __restrict__ dx, int B, int S, int D, int b_per_block,
                                      __nv_bfloat16* __restrict__ dpatch, float* __restrict__ dpos, float* __restrict__ dcls) {
  const int s = blockIdx.x;
  const int b0 = blockIdx.y * b_per_block;
  const int b1 = min(B, b0 + b_per_block);
  const int N = S - 1;
  for (int d = threadIdx.x * 4; d < D; d += blockDim.x * 4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = b0; b < b1; ++b) {
      const float4 f = *reinterpret_cast<const float4*>(dx + (static_cast<int64_t>(b) * S + s) * D + d);
      acc.x += f.x; acc.y += f.y; acc.z += f.z; acc.w += f.w;
      if (s > 0 && dpatch) {
        uint2 u; u.x = pack_bf16(f.x, f.y); u.y = pack_bf16(f.z, f.w);
        *reinterpret_cast<uint2*>(dpatch + (static_cast<int64_t>(b) * N + s - 1) * D + d) = u;
      }
    }
    float* o = dpos + static_cast<int64_t>(s) * D + d;
    atomicAdd(o, acc.x); atomicAdd(o + 1, acc.y); atomicAdd(o + 2, acc.z); atomicAdd(o + 3, acc.w);
    if (s == 0 && dcls) {
      atomicAdd(dcls + d, acc.x); atomicAdd(dcls + d + 1, acc.y); atomicAdd(dcls + d + 2, acc.z); atomicAdd(dcls + d + 3, acc.w);
    }
  }
}

int vit_tokens_bwd_impl(Ctx* ctx, const float* dx, int B, int N, int D, void* dpatch, float* dpos, float* dcls,
                        cudaStream_t st) {
  SIMSEG_CHECK_ARG(B > 0 && N > 0 && D % 4 == 0, "vit_tokens_bwd: bad shape");
  const int S = N + 1;
  int splits = static_cast<int>(cdiv(static_cast<int64_t>(ctx->num_sms) * 8, S));
  if (splits > B) splits = B;
  if (splits < 1) splits = 1;
  const int bpb = static_cast<int>(cdiv(B, splits));
  dim3 grid(S, static_cast<unsigned>(cdiv(B, bpb)));
  const int threads = D / 4 >= 256 ? 256 : ((D / 4 + 31) / 32) * 32;
  vit_tokens_bwd_kernel<<<grid, threads, 0, st>>>(dx, B, S, D, bpb, reinterpret_cast<__nv_bfloat16*>(dpatch), dpos, dcls);
  ctx->launches++;
  SIMSEG_LAUNCH_CHECK();
  return SIMSEG_OK;
}

// ------------------------------------------------------------------------------------------------
__global__ void bert_embed_fwd_kernel(const int64_t* __restrict__ ids, const float* __restrict__ word,
                                      const float* __restrict__ pos, const float* __restrict__ type0, int64_t BT, int T,
                                      int D, float* __restrict__ e) {
  const int D4 = D / 4;
  const int64_t total = BT * D4;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int d = static_cast<int>(i % D4) * 4;
    const int64_t bt = i / D4;
    const int t = static_cast<int>(bt % T);
    const int64_t id = ids[bt];
    const float4 w = *reinterpret_cast<const float4*>(word + id * D + d);
    const float4 p = *reinterpret_cast<const float4*>(pos + static_cast<int64_t>(t) * D + d);
    const float4 y = *reinterpret_cast<const float4*>(type0 + d);
    *reinterpret_cast<float4*>(e + bt * D + d) = make_float4(w.x + y.x + p.x, w.y + y.y + p.y, w.z + y.z + p.z, w.w + y.w + p.w);
  }
}

int bert_embed_fwd_impl(Ctx* ctx, const int64_t* ids, const float* word, const float* pos, const float* type0, int B,
                        int T, int D, float* e, cudaStream_t st) {
  SIMSEG_CHECK_ARG(B > 0 && T > 0 && D % 4 == 0, "bert_embed_fwd: bad shape");
  const int64_t total = static_cast<int64_t>(B) * T * (D / 4);
  const int grid = static_cast<int>(imin64(cdiv(total, 256), static_cast<int64_t>(ctx->num_sms) * 16));
  bert_embed_fwd_kernel<<<grid, 256, 0, st>>>(ids, word, pos, type0, static_cast<int64_t>(B) * T, T, D, e);
  ctx->launches++;
  SIMSEG_LAUNCH_CHECK();
  return SIMSEG_OK;
}

// dword[ids[b,t]] += de[b,t] (vector reductions); dpos[t] += sum_b de[b,t]; dtype0 += sum_{b,t} de[b,t].
// grid = (T, batch_splits), block = D/4 lanes: the batch sum is kept in registers, word rows get red.v4.
__global__ void bert_embed_bwd_kernel(const int64_t* __restrict__ ids, const float* __restrict__ de, int B, int T, int D,
                                      int b_per_block, float* __restrict__ dword, float* __restrict__ dpos,
                                      float* __restrict__ dtype0) {
  const int t = blockIdx.x;
  const int b0 = blockIdx.y * b_per_block;
  const int b1 = min(B, b0 + b_per_block);
  for (int d = threadIdx.x * 4; d < D; d += blockDim.x * 4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = b0; b < b1; ++b) {
      const int64_t bt = static_cast<int64_t>(b) * T + t;
      const float4 f = *reinterpret_cast<const float4*>(de + bt * D + d);
      acc.x += f.x; acc.y += f.y; acc.z += f.z; acc.w += f.w;
      float* w = dword + ids[bt] * D + d;
      asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(w), "f"(f.x), "f"(f.y), "f"(f.z), "f"(f.w) : "memory");
    }
    float* o = dpos + static_cast<int64_t>(t) * D + d;
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(o), "f"(acc.x), "f"(acc.y), "f"(acc.z), "f"(acc.w) : "memory");
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dtype0 + d), "f"(acc.x), "f"(acc.y), "f"(acc.z), "f"(acc.w) : "memory");
  }
}

int bert_embed_bwd_impl(Ctx* ctx, const int64_t* ids, const float* de, int B, int T, int D, float* dword, float* dpos,
                        float* dtype0, cudaStream_t st) {
  SIMSEG_CHECK_ARG(B > 0 && T > 0 && D % 4 == 0, "bert_embed_bwd: bad shape");
  int splits = static_cast<int>(cdiv(static_cast<int64_t>(ctx->num_sms) * 8, T));
  if (splits > B) splits = B;
  if (splits < 1) splits = 1;
  const int bpb = static_cast<int>(cdiv(B, splits));
  dim3 grid(T, static_cast<unsigned>(cdiv(B, bpb)));
  const int threads = D / 4 >= 256 ? 256 : ((D / 4 + 31) / 32) * 32;
  bert_embed_bwd_kernel<<<grid, threads, 0, st>>>(ids, de, B, T, D, bpb, dword, dpos, dtype0);
  ctx->launches++;
  SIMSEG_LAUNCH_CHECK();
  return SIMSEG_OK;
}

// ------------------------------------------------------------------------------------------------
// LoDA head: per (sample, channel) top-k over tokens -> mean -> L2norm.  Block = one sample, one thread per
// channel (reads of x[b,t,:] are coalesced across channels).  k <= 8, kept as a sorted register list.
constexpr int kMaxK = 8;

// K is a template parameter: with a run-time k the sorted list is indexed dynamically and lives in local memory
// (measured: 1.1-1.5 ms for 4096 x 196 x 512; in registers the kernel is a stream of coalesced loads).
template <bool XBF16, int K>
__global__ void topk_pool_l2norm_fwd_kernel(const void* __restrict__ x, int S, int E, int tok_begin, int ntok,
                                            const int64_t* __restrict__ mask, int mask_ld, float eps,
                                            float* __restrict__ pooled, float* __restrict__ emb, int32_t* __restrict__ sel_idx) {
  const int b = blockIdx.x;
  __shared__ float warp_ss[32];
  float ssq = 0.f;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {     // blockDim.x >= E in practice: one pass
    float tv[K];
    int ti[K];
#pragma unroll
    for (int j = 0; j < K; ++j) { tv[j] = -INFINITY; ti[j] = -1; }
    auto insert = [&](float v, int s) {
      // descending list; strict > keeps the earliest token ahead on ties, like a stable top-k.  Once the new value has
      // displaced an entry, that entry travels down and must also pass entries EQUAL to it (they came later) — with a
      // strict compare there the order among equal values flipped and the earlier token could be the one dropped
      if (v > tv[K - 1]) {
        float cur = v;
        int ci = s;
        bool moved = false;
#pragma unroll
        for (int j = 0; j < K; ++j) {
          const bool gt = moved ? (cur >= tv[j]) : (cur > tv[j]);
          moved = moved || gt;
          const float t0 = tv[j];
          const int i0 = ti[j];
          tv[j] = gt ? cur : t0; ti[j] = gt ? ci : i0;
          cur = gt ? t0 : cur; ci = gt ? i0 : ci;
        }
      }
    };
    auto load = [&](int s) {
      float v;
      if (XBF16) v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[(static_cast<int64_t>(b) * S + s) * E + e]);
      else v = reinterpret_cast<const float*>(x)[(static_cast<int64_t>(b) * S + s) * E + e];
      if (mask && mask[static_cast<int64_t>(b) * mask_ld + s] == 0) v = -10000.0f;       // pooling.py:60
      return v;
    };
    int t = 0;
    for (; t + 8 <= ntok; t += 8) {                         // eight independent loads in flight per thread
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = load(tok_begin + t + u);
#pragma unroll
      for (int u = 0; u < 8; ++u) insert(v[u], tok_begin + t + u);
    }
    for (; t < ntok; ++t) insert(load(tok_begin + t), tok_begin + t);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < K; ++j) sum += tv[j];
    const float p = sum / static_cast<float>(K);
    pooled[static_cast<int64_t>(b) * E + e] = p;
    if (sel_idx) {
#pragma unroll
      for (int j = 0; j < K; ++j) sel_idx[(static_cast<int64_t>(b) * K + j) * E + e] = ti[j];
    }
    ssq += p * p;
  }
  if (emb == nullptr) return;
  ssq = warp_sum(ssq);
  if ((threadIdx.x & 31) == 0) warp_ss[threadIdx.x >> 5] = ssq;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < (blockDim.x >> 5); ++w) tot += warp_ss[w];
  const float inv = 1.0f / (sqrtf(tot) + eps);                        // normalization.py:9-10 (eps added)
  for (int e = threadIdx.x; e < E; e += blockDim.x)
    emb[static_cast<int64_t>(b) * E + e] = pooled[static_cast<int64_t>(b) * E + e] * inv;
}

template <int K>
static void launch_topk(bool bf16, int B, int threads, cudaStream_t st, const void* x, int S, int E, int tok_begin, int ntok,
                        const int64_t* mask, int mask_ld, float eps, float* pooled, float* emb, int32_t* sel_idx) {
  if (bf16) topk_pool_l2norm_fwd_kernel<true, K><<<B, threads, 0, st>>>(x, S, E, tok_begin, ntok, mask, mask_ld, eps, pooled, emb, sel_idx);
  else topk_pool_l2norm_fwd_kernel<false, K><<<B, threads, 0, st>>>(x, S, E, tok_begin, ntok, mask, mask_ld, eps, pooled, emb, sel_idx);
}

int topk_pool_l2norm_fwd_impl(Ctx* ctx, const void* x, int x_dtype, int B, int S, int E, int tok_begin, int ntok, int k,
                              const int64_t* mask, int mask_ld, float eps, float* pooled, float* emb, int32_t* sel_idx,
                              cudaStream_t st) {
  SIMSEG_CHECK_ARG(B > 0 && E > 0 && ntok > 0 && tok_begin >= 0 && tok_begin + ntok <= S, "topk_pool: bad token range");
  SIMSEG_CHECK_ARG(k >= 1 && k <= kMaxK && k <= ntok, "topk_pool: k=%d unsupported (1..8, <= ntok)", k);
  SIMSEG_CHECK_ARG(pooled != nullptr, "topk_pool: pooled output required");
  int threads = ((E + 31) / 32) * 32;
  if (threads > 1024) threads = 1024;
  const bool bf = x_dtype == SIMSEG_BF16;
#define TOPK_CASE(KK) case KK: launch_topk<KK>(bf, B, threads, st, x, S, E, tok_begin, ntok, mask, mask_ld, eps, pooled, emb, sel_idx); break
  switch (k) {
    TOPK_CASE(1); TOPK_CASE(2); TOPK_CASE(3); TOPK_CASE(4); TOPK_CASE(5); TOPK_CASE(6); TOPK_CASE(7); TOPK_CASE(8);
  }
#undef TOPK_CASE
  ctx->launches++;
  SIMSEG_LAUNCH_CHECK();
  return SIMSEG_OK;
}

// backward: demb -> dpooled (L2norm backward) -> dense bf16 dx with k non-zeros per (sample, channel).
__global__ void topk_pool_l2norm_bwd_kernel(const float* __restrict__ demb, const float* __restrict__ pooled,
                                            const int32_t* __restrict__ sel_idx, int S, int E, int k, float eps,
                                            int has_l2norm, __nv_bfloat16* __restrict__ dx) {
  const int b = blockIdx.x;
  __shared__ float warp_a[32], warp_b[32];
  // zero this sample's slab (S*E bf16), 16 B per store
  {
    uint4* z = reinterpret_cast<uint4*>(dx + static_cast<int64_t>(b) * S * E);
    const int n16 = (S * E) / 8;
    for (int i = threadIdx.x; i < n16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
  }
  float ss = 0.f, dot = 0.f;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    const float p = pooled[static_cast<int64_t>(b) * E + e];
    ss += p * p;
    dot += p * demb[static_cast<int64_t>(b) * E + e];
  }
  ss = warp_sum(ss); dot = warp_sum(dot);
  if ((threadIdx.x & 31) == 0) { warp_a[threadIdx.x >> 5] = ss; warp_b[threadIdx.x >> 5] = dot; }
  __syncthreads();                                   // also orders the zero-fill before the scatter below
  float tss = 0.f, tdot = 0.f;
  for (int w = 0; w < (blockDim.x >> 5); ++w) { tss += warp_a[w]; tdot += warp_b[w]; }
  const float n = sqrtf(tss);
  const float inv = 1.0f / (n + eps);
  // y = p/(n+eps): dp = dy/(n+eps) - p * (dy.p) / (n (n+eps)^2)
  const float coef = (n > 0.f) ? tdot * inv * inv / n : 0.f;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    const float g = demb[static_cast<int64_t>(b) * E + e];
    float dp = g;
    if (has_l2norm) dp = g * inv - pooled[static_cast<int64_t>(b) * E + e] * coef;
    const __nv_bfloat16 v = __float2bfloat16(dp / static_cast<float>(k));
    for (int j = 0; j < k; ++j) {
      const int s = sel_idx[(static_cast<int64_t>(b) * k + j) * E + e];
      if (s >= 0) dx[(static_cast<int64_t>(b) * S + s) * E + e] = v;
    }
  }
}

int topk_pool_l2norm_bwd_impl(Ctx* ctx, const float* demb, const float* pooled, const int32_t* sel_idx, int B, int S,
                              int E, int k, float eps, int has_l2norm, void* dx, cudaStream_t st) {
  SIMSEG_CHECK_ARG(B > 0 && (S * E) % 8 == 0, "topk_pool_bwd: S*E must be a multiple of 8");
  SIMSEG_CHECK_ARG(k >= 1 && k <= kMaxK, "topk_pool_bwd: k=%d unsupported", k);
  int threads = ((E + 31) / 32) * 32;
  if (threads > 1024) threads = 1024;
  topk_pool_l2norm_bwd_kernel<<<B, threads, 0, st>>>(demb, pooled, sel_idx, S, E, k, eps, has_l2norm,
                                                    reinterpret_cast<__nv_bfloat16*>(dx));
  ctx->launches++;
  SIMSEG_LAUNCH_CHECK();
  return SIMSEG_OK;
}

}  // namespace simseg
