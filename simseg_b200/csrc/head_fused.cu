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

constexpr int kHfThreads = 192;
constexpr int kHfStages = 4;
constexpr int kHfABytes = 128 * 128;            // 128 rows x 64 bf16
constexpr int kHfBBytes = 256 * 128;            // up to 256 rows x 64 bf16
constexpr int kHfStageBytes = kHfABytes + kHfBBytes;
constexpr int kHfSmem = kHfStages * kHfStageBytes + 1024 + 256;

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

template <int K>
__global__ void __launch_bounds__(kHfThreads, 1)
proj_topk_fwd_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_x, const HeadFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kHfStages * kHfStageBytes);
  uint64_t* full_bar = bars;            // [4]
  uint64_t* empty_bar = bars + 4;       // [4]
  uint64_t* acc_full = bars + 8;        // [2]
  uint64_t* acc_empty = bars + 10;      // [2]  (4 epilogue warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_w);
    prefetch_tmap(&tm_x);
    for (int s = 0; s < kHfStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
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
    const int quarter = warp & 3;
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
        float tv[K];
        int ti[K];
#pragma unroll
        for (int j = 0; j < K; ++j) { tv[j] = -INFINITY; ti[j] = -1; }
        for (int ch = c_lo >> 5; ch <= (c_hi - 1) >> 5; ++ch) {
          uint32_t r[32];
          tmem_ld_32x32(t_row + ch * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = ch * 32 + j;
            if (col >= c_lo && col < c_hi) {                          // warp-uniform
              const int s = col - i * p.S;
              float v = __bfloat162float(__float2bfloat16(__uint_as_float(r[j])));
              if (p.mask != nullptr && __ldg(p.mask + static_cast<int64_t>(b) * p.mask_ld + s) == 0) v = -10000.0f;   // pooling.py:60
              // descending list; strict > keeps the earliest token ahead on ties, like a stable top-k
              if (v > tv[K - 1]) {
                float cur = v;
                int ci = s;
#pragma unroll
                for (int q = 0; q < K; ++q) {
                  const bool gt = cur > tv[q];
                  const float t0 = tv[q];
                  const int i0 = ti[q];
                  tv[q] = gt ? cur : t0; ti[q] = gt ? ci : i0;
                  cur = gt ? t0 : cur; ci = gt ? i0 : ci;
                }
              }
            }
          }
        }
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < K; ++j) sum += tv[j];
        p.pooled[static_cast<int64_t>(b) * p.E + e] = sum / static_cast<float>(K);
        if (p.sel_idx != nullptr) {
#pragma unroll
          for (int j = 0; j < K; ++j) p.sel_idx[(static_cast<int64_t>(b) * K + j) * p.E + e] = ti[j];
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

}  // namespace simseg
