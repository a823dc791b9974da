// LoDA head, fused: projection GEMM whose epilogue IS the per-channel top-k pooling, and the backward of
// the pair as two tensor-core GEMMs whose sparse operand is generated in shared memory — the (B, S, E) projection never
// exists in HBM in either direction.
//
//   reference:  image_pool(image_projection(feat))  simseg/models/pipelines/clip.py:87-93 (text: :111-120)
//               SimpleProjection.forward            simseg/models/components/projection.py:45-46   y = x @ W^T  (no bias)
//               TopKPooling.forward                 simseg/models/components/pooling.py:57-65      mean of the k largest tokens,
//                                                   per (sample, channel); masked tokens are set to -10000 first
//
// Forward.  The GEMM is run TRANSPOSED: M = 128 projection channels (TMEM lanes), N = the tokens of one sample (or of `ipt`
// whole samples when they are short: 10 captions of 25 tokens), K = D.  A thread of the epilogue then owns ONE channel and
// walks the token columns of its TMEM lane: the top-k list (value, token) lives in registers, insertion is thread-local, no
// cross-thread reduction exists.  Values are rounded to bf16 before they are compared — what the reference's autocast
// Linear hands to topk — with strict ">" so the earliest token wins ties.
//   warp 0  TMA producer: W tile [128 ch x 64 k] and X tile [N tokens x 64 k] per k-block into a 4-stage ring
//   warp 1  MMA issuer:   tcgen05.mma kind::f16, 128 x N x 16, two accumulator stages of 256 TMEM columns
//   warps 2..5  epilogue: one warp per TMEM lane quarter
//
// Backward.  dY = dL/d(projection) has k non-zeros per (sample, channel): dY[b, s, e] = g[b, e] / k for s in sel[b, :, e].
//   dgrad  dX_b^T [D x tokens] = Wt [D x E] . dY_b^T      M = 128 feature columns, N = tokens of one sample, K = E
//   wgrad  dW [E x D]        += dY_b^T [E x tokens] . X_b  M = 128 channels, N = D (<= 384 per pass), K = tokens, all samples
// In both the dY operand is a [rows][64] K-major SWIZZLE_128B tile that four "generator" warps build in the ring stage:
// scatter the k values of each (sample, channel), hand the stage to the MMA issuer, and — once the MMAs have read it —
// write zeros back to the same few positions (the stage is zero-filled once, at kernel start).
#include "common.cuh"
#include "sm100.cuh"

namespace simseg {

using namespace sm100;

constexpr int kHfGroups = 4;                    // epilogue warps per TMEM lane quarter (column groups)
constexpr int kHfEpiWarps = 4 * kHfGroups;
constexpr int kHfThreads = 32 * (2 + kHfEpiWarps);
constexpr int kHfStages = 4;
constexpr int kHfABytes = 128 * 128;            // 128 rows x 64 bf16
constexpr int kHfBBytes = 256 * 128;            // up to 256 rows x 64 bf16
constexpr int kHfStageBytes = kHfABytes + kHfBBytes;
constexpr int kHfMaxLists = 8;                  // samples per tile x k <= 8 (partial top-k lists exchanged through smem)
constexpr int kHfPartBytes = kHfGroups * 4 * kHfMaxLists * 32 * 4;     // [group][quarter][entry][lane] keys
constexpr int kHfSmem = kHfStages * kHfStageBytes + kHfPartBytes + 1024 + 256;

struct HeadFwdParams {
  int32_t B, S, D, E;
  int32_t tok_begin, ntok;
  int32_t ipt;                 // samples per N tile
  int32_t ncols;               // MMA N = ceil16(ipt * S) <= 256
  int32_t n_tiles, m_tiles, kb;
  const int64_t* mask;
  int32_t mask_ld;
  float* pooled;               // [B, E]
  int32_t* sel_idx;            // [B, K, E] or null
};

// A (value, token) pair as ONE unsigned key: the bf16 value, made order-preserving, in the high half and 0xffff - token in the
// low half.  max() on keys = larger value first, earlier token first among equal values — the order of a stable top-k — and
// the sorted insertion is a chain of k (max, min) pairs with no branch and no divergence.  (The first version compared floats
// and inserted under `if (v > smallest)`: with 32 lanes almost every column had SOME lane inserting, the dependent select
// chain ran ~150 clocks per column and the epilogue took 12x the tile's MMA time.)
__device__ __forceinline__ uint32_t topk_key(uint32_t bf16_hi_bits, int token) {
  uint32_t x = bf16_hi_bits;
  x ^= static_cast<uint32_t>(static_cast<int32_t>(x) >> 31) | 0x80000000u;
  return (x & 0xffff0000u) | (0xffffu - static_cast<uint32_t>(token));
}
__device__ __forceinline__ float topk_key_value(uint32_t key) {
  const uint32_t x = key & 0xffff0000u;
  return __uint_as_float((x & 0x80000000u) ? (x ^ 0x80000000u) : (~x & 0xffff0000u));
}
template <int K>
__device__ __forceinline__ void topk_insert(uint32_t (&t)[K], uint32_t x) {
#pragma unroll
  for (int q = 0; q < K; ++q) {
    const uint32_t hi = max(t[q], x);
    x = min(t[q], x);
    t[q] = hi;
  }
}

template <int K>
__global__ void __launch_bounds__(kHfThreads, 1)
proj_topk_fwd_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_x, const HeadFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint32_t* part = reinterpret_cast<uint32_t*>(smem + kHfStages * kHfStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kHfStages * kHfStageBytes + kHfPartBytes);
  uint64_t* full_bar = bars;            // [4]
  uint64_t* empty_bar = bars + 4;       // [4]
  uint64_t* acc_full = bars + 8;        // [2]
  uint64_t* acc_empty = bars + 10;      // [2]  (all epilogue warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_w);
    prefetch_tmap(&tm_x);
    for (int s = 0; s < kHfStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], kHfEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int units = p.n_tiles * p.m_tiles;
  const uint32_t stage_tx = kHfABytes + static_cast<uint32_t>(p.ncols) * 128u;

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      const int tile = u / p.m_tiles, mt = u - tile * p.m_tiles;
      for (int kb = 0; kb < p.kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + stage * kHfStageBytes;
          mbar_arrive_expect_tx(&full_bar[stage], stage_tx);
          tma_load_2d(sa, &tm_w, &full_bar[stage], kb * 64, mt * 128);
          tma_load_2d(sa + kHfABytes, &tm_x, &full_bar[stage], kb * 64, tile * p.ipt * p.S);
        }
        __syncwarp();
        if (++stage == kHfStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(1u, 0u, 0u, 128, static_cast<uint32_t>(p.ncols));
    const uint64_t adesc0 = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
    const uint64_t bdesc0 = make_smem_desc_sw128(smem_u32(smem) + kHfABytes, 16, 1024);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      mbar_wait(&acc_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 256;
      for (int kb = 0; kb < p.kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t ad = adesc0 + static_cast<uint64_t>(stage * (kHfStageBytes >> 4));
          const uint64_t bd = bdesc0 + static_cast<uint64_t>(stage * (kHfStageBytes >> 4));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_f16(d_tmem, ad + 2 * kk, bd + 2 * kk, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (kb == p.kb - 1) umma_commit(&acc_full[acc]);
        }
        __syncwarp();
        if (++stage == kHfStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // =============================== epilogue: 16 warps, warp (quarter, grp) ===============================
    // thread = one channel (TMEM lane); the four warps of a quarter take the 32-column chunks ch = grp (mod 4) of every
    // sample, keep a partial top-k list per sample, and the lists meet in shared memory
    const int quarter = warp & 3;
    const int grp = (warp - 2) >> 2;
    uint32_t* mine = part + ((grp * 4 + quarter) * kHfMaxLists) * 32 + lane;            // [entry][lane]
    const uint32_t* others = part + (quarter * kHfMaxLists) * 32 + lane;                // + g * 4 * kHfMaxLists * 32
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      const int tile = u / p.m_tiles, mt = u - tile * p.m_tiles;
      const int e = mt * 128 + quarter * 32 + lane;
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + acc * 256 + (static_cast<uint32_t>(quarter * 32) << 16);
      for (int i = 0; i < p.ipt; ++i) {
        const int b = tile * p.ipt + i;
        if (b >= p.B) break;
        const int c_lo = i * p.S + p.tok_begin, c_hi = c_lo + p.ntok;
        uint32_t t[K];
#pragma unroll
        for (int q = 0; q < K; ++q) t[q] = 0u;
        for (int ch = c_lo >> 5; ch <= (c_hi - 1) >> 5; ++ch) {
          if ((ch & (kHfGroups - 1)) != grp) continue;                  // warp-uniform
          uint32_t r[32];
          tmem_ld_32x32(t_row + ch * 32, r);
          tmem_ld_wait();
          const bool whole = (ch * 32 >= c_lo) && (ch * 32 + 32 <= c_hi) && p.mask == nullptr;
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const uint32_t pk = pack_bf16(__uint_as_float(r[j]), __uint_as_float(r[j + 1]));   // what autocast's Linear emits
            const int col = ch * 32 + j;
            uint32_t v0 = pk << 16, v1 = pk & 0xffff0000u;
            if (whole) {
              topk_insert<K>(t, topk_key(v0, col - i * p.S));
              topk_insert<K>(t, topk_key(v1, col + 1 - i * p.S));
            } else {
              if (col >= c_lo && col < c_hi) {
                const int s = col - i * p.S;
                if (p.mask != nullptr && __ldg(p.mask + static_cast<int64_t>(b) * p.mask_ld + s) == 0) v0 = 0xc61c0000u;   // bf16(-10000), pooling.py:60
                topk_insert<K>(t, topk_key(v0, s));
              }
              if (col + 1 >= c_lo && col + 1 < c_hi) {
                const int s = col + 1 - i * p.S;
                if (p.mask != nullptr && __ldg(p.mask + static_cast<int64_t>(b) * p.mask_ld + s) == 0) v1 = 0xc61c0000u;
                topk_insert<K>(t, topk_key(v1, s));
              }
            }
          }
        }
#pragma unroll
        for (int q = 0; q < K; ++q) mine[(i * K + q) * 32] = t[q];
      }
      // every TMEM read of this warp is done: hand the accumulator back before the merge
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
      asm volatile("bar.sync %0, %1;" ::"r"(1 + quarter), "r"(32 * kHfGroups) : "memory");
      for (int i = grp; i < p.ipt; i += kHfGroups) {                    // group g merges the samples i = g (mod 4)
        const int b = tile * p.ipt + i;
        if (b >= p.B) break;
        uint32_t t[K];
#pragma unroll
        for (int q = 0; q < K; ++q) t[q] = others[(i * K + q) * 32];
#pragma unroll
        for (int g2 = 1; g2 < kHfGroups; ++g2) {
#pragma unroll
          for (int q = 0; q < K; ++q) topk_insert<K>(t, others[(g2 * 4 * kHfMaxLists + i * K + q) * 32]);
        }
        float sum = 0.f;
#pragma unroll
        for (int q = 0; q < K; ++q) {
          const int tok = (t[q] == 0u) ? -1 : static_cast<int>(0xffffu - (t[q] & 0xffffu));
          float v = (t[q] == 0u) ? -INFINITY : topk_key_value(t[q]);
          if (p.mask != nullptr && tok >= 0 && __ldg(p.mask + static_cast<int64_t>(b) * p.mask_ld + tok) == 0) v = -10000.0f;
          sum += v;
          if (p.sel_idx != nullptr) p.sel_idx[(static_cast<int64_t>(b) * K + q) * p.E + e] = tok;
        }
        p.pooled[static_cast<int64_t>(b) * p.E + e] = sum / static_cast<float>(K);
      }
      asm volatile("bar.sync %0, %1;" ::"r"(1 + quarter), "r"(32 * kHfGroups) : "memory");   // lists read: may be rewritten
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// emb[b, :] = pooled[b, :] / (||pooled[b, :]|| + eps)      (normalization.py:6-11); one warp per row
__global__ void l2norm_rows_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int E, float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* xr = x + static_cast<int64_t>(row) * E;
  float ss = 0.f;
  for (int e = lane; e < E; e += 32) { const float v = xr[e]; ss += v * v; }
  ss = warp_sum(ss);
  const float inv = 1.0f / (sqrtf(ss) + eps);
  for (int e = lane; e < E; e += 32) y[static_cast<int64_t>(row) * E + e] = xr[e] * inv;
}

template <int K>
static int launch_head_fwd(const CUtensorMap& tw, const CUtensorMap& tx, const HeadFwdParams& p, int grid, cudaStream_t st) {
  static bool attr_set = false;
  auto kfn = proj_topk_fwd_kernel<K>;
  if (!attr_set) {
    SIMSEG_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, kHfSmem));
    attr_set = true;
  }
  kfn<<<grid, kHfThreads, kHfSmem, st>>>(tw, tx, p);
  return SIMSEG_OK;
}

int proj_topk_fwd_impl(Ctx* ctx, const void* x, const void* w, int B, int S, int D, int E, int tok_begin, int ntok, int k,
                       const int64_t* mask, int mask_ld, float eps, float* pooled, float* emb, int32_t* sel_idx,
                       cudaStream_t st) {
  SIMSEG_CHECK_ARG(B > 0 && S > 0 && S <= 256, "proj_topk: S=%d unsupported (1..256 tokens per sample)", S);
  SIMSEG_CHECK_ARG(D % 64 == 0 && E % 128 == 0, "proj_topk: D=%d must be a multiple of 64 and E=%d of 128", D, E);
  SIMSEG_CHECK_ARG(ntok > 0 && tok_begin >= 0 && tok_begin + ntok <= S, "proj_topk: bad token range");
  SIMSEG_CHECK_ARG(k >= 1 && k <= 8 && k <= ntok, "proj_topk: k=%d unsupported (1..8, <= ntok)", k);
  SIMSEG_CHECK_ARG(pooled != nullptr, "proj_topk: pooled output required");
  SIMSEG_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0, "proj_topk: x / w must be 16-byte aligned");
  HeadFwdParams p{};
  p.B = B; p.S = S; p.D = D; p.E = E; p.tok_begin = tok_begin; p.ntok = ntok;
  p.ipt = 256 / S;
  if (p.ipt * k > kHfMaxLists) p.ipt = kHfMaxLists / k;      // partial lists of a tile travel through shared memory
  if (p.ipt < 1) p.ipt = 1;
  p.ncols = (p.ipt * S + 15) & ~15;
  p.n_tiles = static_cast<int>(cdiv(B, p.ipt));
  p.m_tiles = E / 128;
  p.kb = D / 64;
  p.mask = mask; p.mask_ld = mask_ld;
  p.pooled = pooled; p.sel_idx = sel_idx;
  CUtensorMap tw, tx;
  int rc;
  if ((rc = make_tmap(&tw, w, 2, E, D, D, 64, 128))) return rc;
  if ((rc = make_tmap(&tx, x, 2, static_cast<int64_t>(B) * S, D, D, 64, p.ncols))) return rc;
  const int units = p.n_tiles * p.m_tiles;
  const int grid = units < ctx->num_sms ? units : ctx->num_sms;
#define HF_CASE(KK) case KK: rc = launch_head_fwd<KK>(tw, tx, p, grid, st); break
  switch (k) { HF_CASE(1); HF_CASE(2); HF_CASE(3); HF_CASE(4); HF_CASE(5); HF_CASE(6); HF_CASE(7); HF_CASE(8); }
#undef HF_CASE
  if (rc) return rc;
  ctx->launches++;
  SIMSEG_LAUNCH_CHECK();
  if (emb != nullptr) {
    l2norm_rows_kernel<<<static_cast<int>(cdiv(B, 8)), 256, 0, st>>>(pooled, emb, B, E, eps);
    ctx->launches++;
    SIMSEG_LAUNCH_CHECK();
  }
  return SIMSEG_OK;
}


// ================================================================================================ backward
// gy[b, e] = dL/dpooled[b, e] / k  from dL/demb (L2norm backward, normalization.py:6-11: y = p / (||p|| + eps))
__global__ void pool_l2norm_bwd_vec_kernel(const float* __restrict__ demb, const float* __restrict__ pooled, int B, int E,
                                           int k, float eps, int has_l2norm, float* __restrict__ gy) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* g = demb + static_cast<int64_t>(row) * E;
  const float* pr = pooled + static_cast<int64_t>(row) * E;
  float ss = 0.f, dot = 0.f;
  for (int e = lane; e < E; e += 32) { const float v = pr[e]; ss += v * v; dot += v * g[e]; }
  ss = warp_sum(ss); dot = warp_sum(dot);
  const float n = sqrtf(ss);
  const float inv = 1.0f / (n + eps);
  const float coef = (n > 0.f) ? dot * inv * inv / n : 0.f;         // dp = dy/(n+eps) - p (dy.p) / (n (n+eps)^2)
  const float rk = 1.0f / static_cast<float>(k);
  for (int e = lane; e < E; e += 32) {
    float dp = g[e];
    if (has_l2norm) dp = g[e] * inv - pr[e] * coef;
    gy[static_cast<int64_t>(row) * E + e] = dp * rk;
  }
}

__device__ __forceinline__ void sts_u16(uint32_t addr, uint16_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}
__device__ __forceinline__ uint16_t bf16_bits(float v) {
  const __nv_bfloat16 h = __float2bfloat16(v);
  return *reinterpret_cast<const uint16_t*>(&h);
}

// ---------------------------------------------------------------------------------------- dgrad
// dX[b, s, d] = sum_e dY[b, s, e] Wt[d, e]       as   acc[d (lane), token (column)] = sum_e Wt[d, e] . dY[token, e]
//   warp 0      TMA: Wt tile [128 d x 64 e] per k-block
//   warp 1      MMA issuer
//   warps 2..5  generators, warp g owns ring stage g: scatter this k-block's non-zeros of dY into the stage's
//               [tokens x 64 e] K-major tile, and take them out again once the MMAs have read the stage
//   warps 6..9  epilogue: fp32 rows of dX, 32 lanes = 32 consecutive d of one token (128-byte stores)
constexpr int kHdThreads = 320;
constexpr int kHdMaxItems = 16;                 // scattered elements per generator lane and stage: 2 x samples per tile x k

struct HeadDgradParams {
  int32_t B, S, D, E, K;
  int32_t ipt, ncols, n_tiles, m_tiles, kb;
  const float* gy;             // [B, E] dL/dpooled / k
  const int32_t* sel;          // [B, K, E] selected tokens (absolute index, -1 = none)
  float* dx;                   // [B * S, D]
};

__global__ void __launch_bounds__(kHdThreads, 1)
proj_topk_dgrad_kernel(const __grid_constant__ CUtensorMap tm_wt, const HeadDgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kHfStages * kHfStageBytes);
  uint64_t* full_bar = bars;            // [4]  TMA (expect_tx) + generator warp
  uint64_t* empty_bar = bars + 4;       // [4]
  uint64_t* acc_full = bars + 8;        // [2]
  uint64_t* acc_empty = bars + 10;      // [2]  (4 epilogue warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_wt);
    for (int s = 0; s < kHfStages; ++s) { mbar_init(&full_bar[s], 2); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  {
    // the generated operand tiles start out all-zero and are kept that way between uses
    for (int s = 0; s < kHfStages; ++s) {
      uint4* z = reinterpret_cast<uint4*>(smem + s * kHfStageBytes + kHfABytes);
      for (int i = threadIdx.x; i < kHfBBytes / 16; i += kHdThreads) z[i] = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int units = p.n_tiles * p.m_tiles;

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      const int mt = u % p.m_tiles;
      for (int kb = 0; kb < p.kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_bar[stage], kHfABytes);
          tma_load_2d(smem + stage * kHfStageBytes, &tm_wt, &full_bar[stage], kb * 64, mt * 128);
        }
        __syncwarp();
        if (++stage == kHfStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(1u, 0u, 0u, 128, static_cast<uint32_t>(p.ncols));
    const uint64_t adesc0 = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
    const uint64_t bdesc0 = make_smem_desc_sw128(smem_u32(smem) + kHfABytes, 16, 1024);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      mbar_wait(&acc_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 256;
      for (int kb = 0; kb < p.kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t ad = adesc0 + static_cast<uint64_t>(stage * (kHfStageBytes >> 4));
          const uint64_t bd = bdesc0 + static_cast<uint64_t>(stage * (kHfStageBytes >> 4));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_f16(d_tmem, ad + 2 * kk, bd + 2 * kk, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (kb == p.kb - 1) umma_commit(&acc_full[acc]);
        }
        __syncwarp();
        if (++stage == kHfStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp < 6) {
    // =============================== generators ===============================
    // Item `it` of a lane = channel c = lane + 32 (it & 1) of the k-block, (sample, rank) pair number it >> 1.  All selected-
    // token and value loads of a stage are issued before the first one is used (one L2 round trip per stage, not one per item).
    const int g = warp - 2;                                   // ring stage owned by this warp
    const uint32_t sB = smem_u32(smem + g * kHfStageBytes + kHfABytes);
    const int pairs = p.ipt * p.K;                            // <= kHdMaxItems / 2
    uint32_t prev[kHdMaxItems];
#pragma unroll
    for (int it = 0; it < kHdMaxItems; ++it) prev[it] = 0xffffffffu;
    for (int n = g;; n += kHfStages) {
      const int ul = n / p.kb;
      const int u = blockIdx.x + ul * gridDim.x;
      if (u >= units) break;
      const int kb = n - ul * p.kb;
      const int tile = u / p.m_tiles;
      int sv[kHdMaxItems];
      float gv[kHdMaxItems];
      {
        int i = 0, j = 0;
#pragma unroll
        for (int it = 0; it < kHdMaxItems; ++it) {
          const int e = kb * 64 + lane + 32 * (it & 1);
          const int b = tile * p.ipt + i;
          sv[it] = -1;
          gv[it] = 0.f;
          if ((it >> 1) < pairs && b < p.B) {
            sv[it] = __ldg(p.sel + (static_cast<int64_t>(b) * p.K + j) * p.E + e);
            gv[it] = __ldg(p.gy + static_cast<int64_t>(b) * p.E + e);
          }
          if (it & 1) { if (++j == p.K) { j = 0; ++i; } }
        }
      }
      mbar_wait(&empty_bar[g], (static_cast<uint32_t>(n / kHfStages) & 1u) ^ 1u);
      // take out what the previous use of this stage put in (every lane, before anything new goes in: a new position of
      // one lane may be an old position of another)
#pragma unroll
      for (int it = 0; it < kHdMaxItems; ++it) {
        if (prev[it] != 0xffffffffu) sts_u16(sB + prev[it], 0);
        prev[it] = 0xffffffffu;
      }
      __syncwarp();
      {
        int i = 0, j = 0;
#pragma unroll
        for (int it = 0; it < kHdMaxItems; ++it) {
          if (sv[it] >= 0) {
            const int c = lane + 32 * (it & 1);
            const int row = i * p.S + sv[it];
            const uint32_t off = static_cast<uint32_t>(row) * 128u + (static_cast<uint32_t>((c >> 3) ^ (row & 7)) << 4) + (c & 7) * 2;
            sts_u16(sB + off, bf16_bits(gv[it]));
            prev[it] = off;
          }
          if (it & 1) { if (++j == p.K) { j = 0; ++i; } }
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[g]);
    }
  } else {
    // =============================== epilogue ===============================
    const int quarter = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      const int tile = u / p.m_tiles, mt = u - tile * p.m_tiles;
      const int d = mt * 128 + quarter * 32 + lane;
      const int64_t row0 = static_cast<int64_t>(tile) * p.ipt * p.S;
      const int64_t rows_total = static_cast<int64_t>(p.B) * p.S;
      const int live = p.ipt * p.S;                            // columns of the tile that are tokens of its samples
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + acc * 256 + (static_cast<uint32_t>(quarter * 32) << 16);
      for (int ch = 0; ch * 32 < live; ++ch) {
        uint32_t r[32];
        tmem_ld_32x32(t_row + ch * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = ch * 32 + j;
          if (col < live && row0 + col < rows_total) p.dx[(row0 + col) * p.D + d] = __uint_as_float(r[j]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---------------------------------------------------------------------------------------- wgrad
// dW[e, d] += sum over ALL token rows r = b * S + s of dY[r, e] X[r, d]      (split over the row axis, red.global.add)
//   acc[e (lane), d (column)]:  A = dY^T tile [128 e x 64 rows] K-major, generated;  B = X rows [64 rows x nw d], MN-major,
//   one 3-D TMA box {64 d, 64 rows, nw / 64 atoms}; nw <= 384 columns = one accumulator, two MMAs (N = 256 + nw - 256) per k-step
//   warp 0 TMA, warp 1 MMA, warps 2..4 generators (one ring stage each), warps 5..8 epilogue
constexpr int kHwThreads = 288;
constexpr int kHwStages = 3;
constexpr int kHwABytes = 128 * 128;
constexpr int kHwBBytes = 6 * 8192;             // up to 6 atoms of [64 rows][128 B]
constexpr int kHwStageBytes = kHwABytes + kHwBBytes;
constexpr int kHwSmem = kHwStages * kHwStageBytes + 1024 + 256;

struct HeadWgradParams {
  int32_t B, S, D, E, K;
  int32_t nw;                  // d columns per unit (128 | 256 | 384)
  int32_t m_tiles, n_pass, splits, kb_total, kb_per_split;
  const float* gy;
  const int32_t* sel;
  float* dw;                   // [E, D] fp32, accumulated into
};

__global__ void __launch_bounds__(kHwThreads, 1)
proj_topk_wgrad_kernel(const __grid_constant__ CUtensorMap tm_x, const HeadWgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kHwStages * kHwStageBytes);
  uint64_t* full_bar = bars;            // [3]  TMA (expect_tx) + generator warp
  uint64_t* empty_bar = bars + 3;       // [3]
  uint64_t* acc_full = bars + 6;
  uint64_t* acc_empty = bars + 7;       // (4 epilogue warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_x);
    for (int s = 0; s < kHwStages; ++s) { mbar_init(&full_bar[s], 2); mbar_init(&empty_bar[s], 1); }
    mbar_init(acc_full, 1); mbar_init(acc_empty, 4);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int units = p.m_tiles * p.n_pass * p.splits;
  const uint32_t b_tx = static_cast<uint32_t>(p.nw / 64) * 8192u;
  // unit -> (split, n pass, channel tile): channel tiles of one (split, pass) are neighbours, they read the same X rows
  auto unit_kb = [&](int u, int& mt, int& np, int& kb0, int& kb1) {
    mt = u % p.m_tiles;
    const int r = u / p.m_tiles;
    np = r % p.n_pass;
    const int sp = r / p.n_pass;
    kb0 = sp * p.kb_per_split;
    kb1 = min(kb0 + p.kb_per_split, p.kb_total);
  };

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      int mt, np, kb0, kb1;
      unit_kb(u, mt, np, kb0, kb1);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_bar[stage], b_tx);
          tma_load_3d(smem + stage * kHwStageBytes + kHwABytes, &tm_x, &full_bar[stage], 0, kb * 64, np * (p.nw / 64));
        }
        __syncwarp();
        if (++stage == kHwStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    const int n1 = p.nw < 256 ? p.nw : 256, n2 = p.nw - n1;
    const uint32_t idesc1 = make_idesc(1u, 0u, 1u, 128, static_cast<uint32_t>(n1));
    const uint32_t idesc2 = make_idesc(1u, 0u, 1u, 128, static_cast<uint32_t>(n2 > 0 ? n2 : 16));
    const uint64_t adesc0 = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
    const uint64_t bdesc0 = make_smem_desc_sw128(smem_u32(smem) + kHwABytes, 8192, 1024);     // MN-major: atoms 8 KB apart
    int stage = 0;
    uint32_t phase = 0, ucount = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++ucount) {
      int mt, np, kb0, kb1;
      unit_kb(u, mt, np, kb0, kb1);
      mbar_wait(acc_empty, (ucount & 1) ^ 1);
      tc_fence_after();
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t ad = adesc0 + static_cast<uint64_t>(stage * (kHwStageBytes >> 4));
          const uint64_t bd = bdesc0 + static_cast<uint64_t>(stage * (kHwStageBytes >> 4));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint32_t accf = (kb > kb0 || kk > 0) ? 1u : 0u;
            umma_f16(tmem_base, ad + 2 * kk, bd + 128 * kk, idesc1, accf);
            if (n2 > 0) umma_f16(tmem_base + 256, ad + 2 * kk, bd + (4 * 8192 >> 4) + 128 * kk, idesc2, accf);
          }
          umma_commit(&empty_bar[stage]);
          if (kb == kb1 - 1) umma_commit(acc_full);
        }
        __syncwarp();
        if (++stage == kHwStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp < 5) {
    // =============================== generators ===============================
    const int g = warp - 2;
    uint8_t* sA = smem + g * kHwStageBytes;
    const uint32_t sA32 = smem_u32(sA);
    // flattened (unit, k-block) iterations of this CTA; warp g takes n = g, g + 3, ...
    int n = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      int mt, np, kb0, kb1;
      unit_kb(u, mt, np, kb0, kb1);
      for (int kb = kb0; kb < kb1; ++kb, ++n) {
        if (n % kHwStages != g) continue;
        mbar_wait(&empty_bar[g], (static_cast<uint32_t>(n / kHwStages) & 1u) ^ 1u);
        uint4* z = reinterpret_cast<uint4*>(sA);
#pragma unroll
        for (int i = 0; i < kHwABytes / 16 / 32; ++i) z[lane + 32 * i] = make_uint4(0, 0, 0, 0);
        __syncwarp();
        const int64_t r0 = static_cast<int64_t>(kb) * 64;
        const int b_lo = static_cast<int>(r0 / p.S);
        int b_hi = static_cast<int>((r0 + 63) / p.S);
        if (b_hi >= p.B) b_hi = p.B - 1;
        for (int b = b_lo; b <= b_hi; ++b) {
          const int64_t base = static_cast<int64_t>(b) * p.S - r0;          // tile row of token 0 of sample b
          // lane -> channels c = lane + 32 m (m = 0..3); all k x 4 selections and the 4 values are loaded before use
          int sv[32];
          float gv[4];
#pragma unroll
          for (int m = 0; m < 4; ++m) gv[m] = __ldg(p.gy + static_cast<int64_t>(b) * p.E + mt * 128 + lane + 32 * m);
#pragma unroll
          for (int t = 0; t < 32; ++t) {
            sv[t] = -1;
            if ((t >> 2) < p.K) sv[t] = __ldg(p.sel + (static_cast<int64_t>(b) * p.K + (t >> 2)) * p.E + mt * 128 + lane + 32 * (t & 3));
          }
#pragma unroll
          for (int t = 0; t < 32; ++t) {
            const int64_t row = base + sv[t];
            if (sv[t] >= 0 && row >= 0 && row < 64) {
              const int c = lane + 32 * (t & 3);
              const int rr = static_cast<int>(row);
              const uint32_t off = static_cast<uint32_t>(c) * 128u + (static_cast<uint32_t>((rr >> 3) ^ (c & 7)) << 4) + (rr & 7) * 2;
              sts_u16(sA32 + off, bf16_bits(gv[t & 3]));
            }
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[g]);
      }
    }
  } else {
    // =============================== epilogue ===============================
    const int quarter = warp & 3;
    uint32_t ucount = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++ucount) {
      int mt, np, kb0, kb1;
      unit_kb(u, mt, np, kb0, kb1);
      const int e = mt * 128 + quarter * 32 + lane;
      mbar_wait(acc_full, ucount & 1);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
      float* dst = p.dw + static_cast<int64_t>(e) * p.D + np * p.nw;
      for (int ch = 0; ch * 32 < p.nw; ++ch) {
        uint32_t r[32];
        tmem_ld_32x32(t_row + ch * 32, r);
        tmem_ld_wait();
        if (kb1 > kb0) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst + ch * 32 + 4 * q), "f"(__uint_as_float(r[4 * q])),
                         "f"(__uint_as_float(r[4 * q + 1])), "f"(__uint_as_float(r[4 * q + 2])), "f"(__uint_as_float(r[4 * q + 3])) : "memory");
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

int proj_topk_bwd_impl(Ctx* ctx, const float* demb, const float* pooled, const int32_t* sel_idx, const void* x, const void* wt,
                       int B, int S, int D, int E, int k, float eps, int has_l2norm, float* gy, float* dx, float* dw,
                       cudaStream_t st) {
  SIMSEG_CHECK_ARG(B > 0 && S > 0 && S <= 256, "proj_topk_bwd: S=%d unsupported (1..256 tokens per sample)", S);
  SIMSEG_CHECK_ARG(D % 128 == 0 && E % 128 == 0, "proj_topk_bwd: D=%d and E=%d must be multiples of 128", D, E);
  SIMSEG_CHECK_ARG(k >= 1 && k <= 8, "proj_topk_bwd: k=%d unsupported", k);
  SIMSEG_CHECK_ARG(demb && pooled && sel_idx && gy, "proj_topk_bwd: demb / pooled / sel_idx / gy required");
  // 1. dL/dpooled / k
  pool_l2norm_bwd_vec_kernel<<<static_cast<int>(cdiv(B, 8)), 256, 0, st>>>(demb, pooled, B, E, k, eps, has_l2norm, gy);
  ctx->launches++;
  SIMSEG_LAUNCH_CHECK();
  int rc;
  // 2. dX = dY W
  if (dx != nullptr) {
    SIMSEG_CHECK_ARG(wt != nullptr && (reinterpret_cast<uintptr_t>(wt) & 15) == 0, "proj_topk_bwd: transposed weight copy [D, E] missing / unaligned");
    HeadDgradParams p{};
    p.B = B; p.S = S; p.D = D; p.E = E; p.K = k;
    int ipt = 256 / S;
    if (ipt * k > kHdMaxItems / 2) ipt = (kHdMaxItems / 2) / k;   // a generator lane keeps 2 x ipt x k scattered positions in registers
    if (ipt < 1) ipt = 1;
    p.ipt = ipt;
    p.ncols = (ipt * S + 15) & ~15;
    p.n_tiles = static_cast<int>(cdiv(B, ipt));
    p.m_tiles = D / 128;
    p.kb = E / 64;
    p.gy = gy; p.sel = sel_idx; p.dx = dx;
    CUtensorMap tw;
    if ((rc = make_tmap(&tw, wt, 2, D, E, E, 64, 128))) return rc;
    static bool attr_set = false;
    if (!attr_set) {
      SIMSEG_CUDA(cudaFuncSetAttribute(proj_topk_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kHfSmem));
      attr_set = true;
    }
    const int units = p.n_tiles * p.m_tiles;
    const int grid = units < ctx->num_sms ? units : ctx->num_sms;
    proj_topk_dgrad_kernel<<<grid, kHdThreads, kHfSmem, st>>>(tw, p);
    ctx->launches++;
    SIMSEG_LAUNCH_CHECK();
  }
  // 3. dW += dY^T X
  if (dw != nullptr) {
    SIMSEG_CHECK_ARG(x != nullptr && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "proj_topk_bwd: x missing / unaligned");
    HeadWgradParams p{};
    p.B = B; p.S = S; p.D = D; p.E = E; p.K = k;
    p.nw = D % 384 == 0 ? 384 : (D % 256 == 0 ? 256 : 128);
    p.m_tiles = E / 128;
    p.n_pass = D / p.nw;
    p.kb_total = static_cast<int>(cdiv(static_cast<int64_t>(B) * S, 64));
    int splits = ctx->num_sms / (p.m_tiles * p.n_pass);
    if (splits < 1) splits = 1;
    if (splits > p.kb_total) splits = p.kb_total;
    p.kb_per_split = static_cast<int>(cdiv(p.kb_total, splits));
    p.splits = static_cast<int>(cdiv(p.kb_total, p.kb_per_split));
    p.gy = gy; p.sel = sel_idx; p.dw = dw;
    CUtensorMap tx;
    if ((rc = make_tmap_mn3d(&tx, x, 2, static_cast<int64_t>(B) * S, D, D, 64, p.nw / 64))) return rc;
    static bool attr_set = false;
    if (!attr_set) {
      SIMSEG_CUDA(cudaFuncSetAttribute(proj_topk_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kHwSmem));
      attr_set = true;
    }
    const int units = p.m_tiles * p.n_pass * p.splits;
    const int grid = units < ctx->num_sms ? units : ctx->num_sms;
    proj_topk_wgrad_kernel<<<grid, kHwThreads, kHwSmem, st>>>(tx, p);
    ctx->launches++;
    SIMSEG_LAUNCH_CHECK();
  }
  return SIMSEG_OK;
}

}  // namespace simseg
