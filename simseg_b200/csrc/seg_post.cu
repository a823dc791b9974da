// Zero-shot segmentation glue around the patch-text map: the steps of tools/seg_evaluation.py that
// sit either side of the similarity kernel, kept on the GPU so that no per-class host round trip is left:
//   class_embed   :57-75   mean over the prompt embeddings of a class, then /= ||.|| (no eps)
//   select        :119-150 image-level class scores -> top-k -> threshold mean + std (unbiased) -> up to 5 candidates
//                          (classes 0 and 255 are skipped, the scan stops at the first score below the threshold)
//   upsample_norm :136-150 per candidate: its column of the map, nearest x16 up-sampling, min-max normalisation
#include "common.cuh"

namespace simseg {

// one block per class: emb[c,:] = mean_p prompt[c,p,:] / || mean ||
__global__ void __launch_bounds__(256) class_embed_kernel(const float* __restrict__ prompt, int P, int E, float* __restrict__ out) {
  const int c = blockIdx.x;
  const float* base = prompt + static_cast<int64_t>(c) * P * E;
  __shared__ float red[8];
  float ss = 0.f;
  // E <= 1024: up to 4 columns per thread, kept in registers
  float m[4] = {0.f, 0.f, 0.f, 0.f};
  const float invP = 1.0f / static_cast<float>(P);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int e = threadIdx.x + k * 256;
    if (e < E) {
      float s = 0.f;
      for (int pidx = 0; pidx < P; ++pidx) s += base[static_cast<int64_t>(pidx) * E + e];
      m[k] = s * invP;
      ss += m[k] * m[k];
    }
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) tot += red[w];
  const float inv = 1.0f / sqrtf(tot);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int e = threadIdx.x + k * 256;
    if (e < E) out[static_cast<int64_t>(c) * E + e] = m[k] * inv;
  }
}

// one warp per image: scores[c] = img[b,:] . text[c,:]; top-k by repeated first-max argmax (torch.topk order for
// distinct scores); threshold = mean + std (unbiased, torch.std default); candidates as in the reference loop
__global__ void __launch_bounds__(32) seg_select_kernel(const float* __restrict__ img, const float* __restrict__ text, int B, int C,
                                                       int E, int topk, int max_cand, float* __restrict__ scores_out,
                                                       int32_t* __restrict__ cand, float* __restrict__ threshold_out) {
  extern __shared__ float s_sc[];                 // [C] scores, then [topk] values + [topk] indices
  const int b = blockIdx.x, lane = threadIdx.x;
  const float* x = img + static_cast<int64_t>(b) * E;
  for (int c = 0; c < C; ++c) {
    float s = 0.f;
    for (int e = lane; e < E; e += 32) s += x[e] * text[static_cast<int64_t>(c) * E + e];
    s = warp_sum(s);
    if (lane == 0) {
      s_sc[c] = s;
      if (scores_out) scores_out[static_cast<int64_t>(b) * C + c] = s;
    }
  }
  __syncwarp();
  float* tv = s_sc + C;
  int* ti = reinterpret_cast<int*>(tv + topk);
  // repeated argmax; taken entries are marked in a bitmask held across lanes (C <= 1024 -> 32 bits per lane)
  uint32_t taken = 0;                              // lane l owns classes l, l+32, ...
  for (int t = 0; t < topk; ++t) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int c = lane, j = 0; c < C; c += 32, ++j) {
      const float v = s_sc[c];
      if (!((taken >> j) & 1u) && (v > best)) { best = v; bi = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if ((bi & 31) == lane && bi < C) taken |= 1u << (bi >> 5);
    if (lane == 0) { tv[t] = best; ti[t] = bi; }
  }
  __syncwarp();
  if (lane == 0) {
    float mean = 0.f;
    for (int t = 0; t < topk; ++t) mean += tv[t];
    mean /= static_cast<float>(topk);
    float var = 0.f;
    for (int t = 0; t < topk; ++t) var += (tv[t] - mean) * (tv[t] - mean);
    const float sd = topk > 1 ? sqrtf(var / static_cast<float>(topk - 1)) : 0.f;
    const float thr = mean + sd;
    if (threshold_out) threshold_out[b] = thr;
    int n = 0;
    bool stop = false;
    const int scan = max_cand < topk ? max_cand : topk;
    for (int t = 0; t < max_cand; ++t) cand[static_cast<int64_t>(b) * max_cand + t] = -1;
    for (int t = 0; t < scan && !stop; ++t) {
      const int idx = ti[t];
      if (idx == 0 || idx == 255) continue;        // seg_evaluation.py:127-128
      if (tv[t] < thr) { stop = true; break; }     // :143-144
      cand[static_cast<int64_t>(b) * max_cand + n++] = idx;
    }
  }
}

// one block per (image, candidate): min / max of the class column over the N patches, then the nearest-up-sampled,
// min-max normalised map (h*scale x w*scale fp32, 128-bit stores)
__global__ void __launch_bounds__(256) seg_upsample_norm_kernel(const float* __restrict__ sim, const int32_t* __restrict__ cand,
                                                               int N, int C, int K, int h, int w, int scale,
                                                               float* __restrict__ out) {
  const int b = blockIdx.x / K, k = blockIdx.x % K;
  const int cls = cand[static_cast<int64_t>(b) * K + k];
  const int W = w * scale, Hh = h * scale;
  float* o = out + static_cast<int64_t>(blockIdx.x) * Hh * W;
  if (cls < 0) {
    for (int i = threadIdx.x; i < Hh * W; i += 256) o[i] = 0.f;
    return;
  }
  const float* col = sim + static_cast<int64_t>(b) * N * C + cls;
  float mn = INFINITY, mx = -INFINITY;
  for (int n = threadIdx.x; n < N; n += 256) {
    const float v = col[static_cast<int64_t>(n) * C];
    mn = fminf(mn, v); mx = fmaxf(mx, v);
  }
  __shared__ float smn[8], smx[8];
  mn = -warp_max(-mn); mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5] = mn; smx[threadIdx.x >> 5] = mx; }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) { mn = fminf(mn, smn[i]); mx = fmaxf(mx, smx[i]); }
  const float inv = 1.0f / (mx - mn);
  for (int i = threadIdx.x; i < Hh * W; i += 256) {
    const int y = i / W, xx = i - y * W;
    const int n = (y / scale) * w + xx / scale;
    o[i] = (col[static_cast<int64_t>(n) * C] - mn) * inv;
  }
}

int seg_class_embed_impl(Ctx* ctx, const float* prompt, int C, int P, int E, float* out, cudaStream_t st) {
  SIMSEG_CHECK_ARG(C > 0 && P > 0 && E > 0 && E <= 1024, "class_embed: C=%d P=%d E=%d unsupported (E <= 1024)", C, P, E);
  class_embed_kernel<<<C, 256, 0, st>>>(prompt, P, E, out);
  ctx->launches++;
  SIMSEG_LAUNCH_CHECK();
  return SIMSEG_OK;
}

int seg_select_impl(Ctx* ctx, const float* img, const float* text, int B, int C, int E, int topk, int max_cand, float* scores,
                    int32_t* cand, float* threshold, cudaStream_t st) {
  SIMSEG_CHECK_ARG(B > 0 && C > 0 && C <= 1024 && E > 0 && topk > 0 && topk <= C && max_cand > 0,
                   "seg_select: B=%d C=%d E=%d topk=%d max_cand=%d unsupported", B, C, E, topk, max_cand);
  const int smem = (C + 2 * topk) * 4;
  seg_select_kernel<<<B, 32, smem, st>>>(img, text, B, C, E, topk, max_cand, scores, cand, threshold);
  ctx->launches++;
  SIMSEG_LAUNCH_CHECK();
  return SIMSEG_OK;
}

int seg_upsample_norm_impl(Ctx* ctx, const float* sim, const int32_t* cand, int B, int N, int C, int K, int h, int w, int scale,
                           float* out, cudaStream_t st) {
  SIMSEG_CHECK_ARG(B > 0 && K > 0 && h * w == N && scale > 0, "seg_upsample_norm: h*w must equal N (h=%d w=%d N=%d)", h, w, N);
  seg_upsample_norm_kernel<<<B * K, 256, 0, st>>>(sim, cand, N, C, K, h, w, scale, out);
  ctx->launches++;
  SIMSEG_LAUNCH_CHECK();
  return SIMSEG_OK;
}

}  // namespace simseg
