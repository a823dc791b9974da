// Zero-shot segmentation glue around the patch-text map: the steps of tools/seg_evaluation.py that
// sit either side of the similarity kernel, kept on the GPU so that no per-class host round trip is left:
//   class_embed   :57-75   mean over the prompt embeddings of a class, then /= ||.|| (no eps)
//   select        :119-150 image-level class scores -> top-k -> threshold mean + std (unbiased) -> up to 5 candidates
//                          (classes 0 and 255 are skipped, the scan stops at the first score below the threshold)
//   upsample_norm :136-150 per candidate: its column of the map, nearest x16 up-sampling, min-max normalisation
//   pos_embed_bicubic :228-230 -> utils/interpolate_pe.py:4-27, the position-embedding resize done at checkpoint load
//                          (224^2 checkpoints evaluated at 288^2: 14x14 -> 18x18 grid)
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

// Bicubic resize of the grid part of a ViT position embedding (utils/interpolate_pe.py:16-22: F.interpolate(mode='bicubic',
// align_corners=False) on the [1,D,g0,g0] view), the extra (class) tokens are copied.  Same arithmetic as torch's
// upsample_bicubic2d: source = (dst+0.5)*g0/g1-0.5 (not clamped), Keys kernel with A=-0.75, border taps clamped.
__device__ __forceinline__ void cubic_coeffs(float t, float (&c)[4]) {
  const float A = -0.75f;
  const float x0 = t + 1.0f, x3 = 2.0f - t, x2 = 1.0f - t;
  c[0] = ((A * x0 - 5.0f * A) * x0 + 8.0f * A) * x0 - 4.0f * A;
  c[1] = ((A + 2.0f) * t - (A + 3.0f)) * t * t + 1.0f;
  c[2] = ((A + 2.0f) * x2 - (A + 3.0f)) * x2 * x2 + 1.0f;
  c[3] = ((A * x3 - 5.0f * A) * x3 + 8.0f * A) * x3 - 4.0f * A;
}
__global__ void __launch_bounds__(128) pos_embed_bicubic_kernel(const float* __restrict__ src, float* __restrict__ dst, int g0, int g1,
                                                               int D, int extra) {
  const int tok = blockIdx.x;
  float* o = dst + static_cast<int64_t>(tok) * D;
  if (tok < extra) {
    for (int d = threadIdx.x; d < D; d += 128) o[d] = src[static_cast<int64_t>(tok) * D + d];
    return;
  }
  const int oy = (tok - extra) / g1, ox = (tok - extra) % g1;
  const float scale = static_cast<float>(g0) / static_cast<float>(g1);
  const float ry = scale * (oy + 0.5f) - 0.5f, rx = scale * (ox + 0.5f) - 0.5f;
  const float fy = floorf(ry), fx = floorf(rx);
  float cy[4], cx[4];
  cubic_coeffs(ry - fy, cy);
  cubic_coeffs(rx - fx, cx);
  const int iy = static_cast<int>(fy), ix = static_cast<int>(fx);
  int64_t row[4];
  int col[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    row[i] = static_cast<int64_t>(extra + min(max(iy - 1 + i, 0), g0 - 1) * g0) * D;
    col[i] = min(max(ix - 1 + i, 0), g0 - 1) * D;
  }
  for (int d = threadIdx.x; d < D; d += 128) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float* r = src + row[i] + d;
      const float v = r[col[0]] * cx[0] + r[col[1]] * cx[1] + r[col[2]] * cx[2] + r[col[3]] * cx[3];
      acc += v * cy[i];
    }
    o[d] = acc;
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

int pos_embed_bicubic_impl(Ctx* ctx, const float* src, float* dst, int g0, int g1, int D, int extra, cudaStream_t st) {
  SIMSEG_CHECK_ARG(g0 > 0 && g1 > 0 && D > 0 && extra >= 0, "pos_embed_bicubic: bad sizes g0=%d g1=%d D=%d extra=%d", g0, g1, D, extra);
  pos_embed_bicubic_kernel<<<extra + g1 * g1, 128, 0, st>>>(src, dst, g0, g1, D, extra);
  ctx->launches++;
  SIMSEG_LAUNCH_CHECK();
  return SIMSEG_OK;
}

}  // namespace simseg
