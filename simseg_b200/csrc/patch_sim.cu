// Fused dense patch-text similarity map (tools/seg_evaluation.py:111-112,136-139 of the reference), bf16 inputs:
//
//   sim[r, c] = (patch[r, :] / max(||patch[r, :]||, 1e-12)) . text[c, :]        argmax[r] = first argmax_c sim[r, c]
//
// ONE pass over HBM: every patch row is read once (TMA -> 128B-swizzled smem ring) and every similarity is written
// once.  The class-text matrix ([C, E] bf16, C <= 256) is loaded once per CTA and stays resident in shared memory as
// the B operand; the contraction runs on tcgen05 (128 x Cpad x 16 MMAs, fp32 accumulators in TMEM, two accumulator
// stages) while four "norm" warps read the same smem stages to accumulate the row sums of squares, so the row
// normalisation costs no extra HBM traffic.  The epilogue scales by 1/||patch||, takes the row argmax, transposes
// 32x32 blocks through padded smem and stores 128-byte row segments (rows of sim are contiguous in HBM).
//
// Persistent CTAs (one per SM, 10 warps):  warp 0 TMA producer | warp 1 MMA issuer | warps 2-5 row norms |
// warps 6-9 epilogue (one TMEM lane quarter each).
// Algorithmic HBM bytes per row: E*2 + C*4 + 4  (DESIGN.md).
#include "common.cuh"
#include "sm100.cuh"

namespace simseg {

using namespace sm100;

constexpr int kPsBM = 128;
constexpr int kPsStageBytes = kPsBM * 128;          // 128 rows x 64 bf16
constexpr int kPsThreads = 320;
constexpr int kPsStagingBytes = 4 * 32 * 33 * 4;
constexpr int kPsMaxStages = 8;

struct PatchSimParams {
  int64_t rows;
  int32_t C, E, npad, kblocks, stages, tiles, normalize;
  uint32_t tmem_cols, acc_stride;
  float* sim;
  int32_t* argmax;
};

__global__ void __launch_bounds__(kPsThreads, 1)
patch_sim_kernel(const __grid_constant__ CUtensorMap tmap_p, const __grid_constant__ CUtensorMap tmap_t, const PatchSimParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int text_kb_bytes = p.npad * 128;
  uint8_t* s_text = smem;
  uint8_t* s_ring = s_text + p.kblocks * text_kb_bytes;
  float* s_stage = reinterpret_cast<float*>(s_ring + p.stages * kPsStageBytes);
  float* s_inv = s_stage + kPsStagingBytes / 4;                       // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_inv + 2 * kPsBM);
  uint64_t* full_bar = bars;                                         // [stages]  TMA -> MMA + norm
  uint64_t* empty_bar = full_bar + kPsMaxStages;                     // [stages]  MMA commit + 4 norm warps -> TMA
  uint64_t* acc_full = empty_bar + kPsMaxStages;                     // [2] MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;                                // [2] epilogue -> MMA, norm
  uint64_t* norm_full = acc_empty + 2;                               // [2] norm -> epilogue
  uint64_t* text_bar = norm_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(text_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_p);
    prefetch_tmap(&tmap_t);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 5);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 4);
      mbar_init(&norm_full[s], 4);
    }
    mbar_init(text_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      mbar_arrive_expect_tx(text_bar, static_cast<uint32_t>(p.kblocks * text_kb_bytes));
      for (int kb = 0; kb < p.kblocks; ++kb)
        tma_load_2d_hint(s_text + kb * text_kb_bytes, &tmap_t, text_bar, kb * 64, 0, kEvictLast);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        const int m0 = tile * kPsBM;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], kPsStageBytes);
          tma_load_2d_hint(s_ring + stage * kPsStageBytes, &tmap_p, &full_bar[stage], kb * 64, m0, kEvictFirst);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(1u, 0u, 0u, kPsBM, static_cast<uint32_t>(p.npad));
      mbar_wait(text_bar, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t s_text_u32 = smem_u32(s_text);
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * p.acc_stride;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(s_ring + stage * kPsStageBytes);
          const uint32_t sb = s_text_u32 + kb * text_kb_bytes;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t adesc = make_smem_desc_sw128(sa + kk * 32, 16, 1024);
            const uint64_t bdesc = make_smem_desc_sw128(sb + kk * 32, 16, 1024);
            umma_f16(d_tmem, adesc, bdesc, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&acc_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp < 6) {
    // =============================== row norms ===============================
    const int row = (warp - 2) * 32 + lane;                 // row inside the tile
    const uint32_t row_off = row * 128;
    const int sw = row & 7;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
      float ss0 = 0.f, ss1 = 0.f;
      for (int kb = 0; kb < p.kblocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        const uint8_t* base = s_ring + stage * kPsStageBytes + row_off;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint4 u = *reinterpret_cast<const uint4*>(base + ((j ^ sw) << 4));
          float a;
          a = bf16_lo(u.x); ss0 = fmaf(a, a, ss0); a = bf16_hi(u.x); ss1 = fmaf(a, a, ss1);
          a = bf16_lo(u.y); ss0 = fmaf(a, a, ss0); a = bf16_hi(u.y); ss1 = fmaf(a, a, ss1);
          a = bf16_lo(u.z); ss0 = fmaf(a, a, ss0); a = bf16_hi(u.z); ss1 = fmaf(a, a, ss1);
          a = bf16_lo(u.w); ss0 = fmaf(a, a, ss0); a = bf16_hi(u.w); ss1 = fmaf(a, a, ss1);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      mbar_wait(&acc_empty[acc], acc_phase ^ 1);             // the epilogue has consumed this slot's previous value
      s_inv[acc * kPsBM + row] = p.normalize ? 1.0f / fmaxf(sqrtf(ss0 + ss1), 1e-12f) : 1.0f;
      __syncwarp();
      if (lane == 0) mbar_arrive(&norm_full[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // =============================== epilogue ===============================
    const int quarter = warp & 3;
    float* stg = s_stage + (warp - 6) * (32 * 33);
    const int row_in_tile = quarter * 32 + lane;
    const int nchunks = (p.npad + 31) >> 5;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
      const int64_t m0 = static_cast<int64_t>(tile) * kPsBM;
      const int64_t wrow0 = m0 + quarter * 32;                // first global row of this warp
      mbar_wait(&norm_full[acc], acc_phase);
      const float inv = s_inv[acc * kPsBM + row_in_tile];
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + acc * p.acc_stride + (static_cast<uint32_t>(quarter * 32) << 16);
      float best = -INFINITY;
      int besti = 0;
      const int64_t rows_left = p.rows - wrow0;
      const int rows_here = rows_left < 32 ? static_cast<int>(rows_left) : 32;      // may be <= 0 for the ragged last tile
      for (int c = 0; c < nchunks; ++c) {
        const int c0 = c * 32;
        uint32_t r[32];
        tmem_ld_32x32(t_row + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float v = __uint_as_float(r[j]) * inv;
          if (c0 + j < p.C && v > best) { best = v; besti = c0 + j; }
          stg[lane * 33 + j] = v;
        }
        __syncwarp();
        if (c0 + lane < p.C) {
          float* out = p.sim + wrow0 * p.C + c0 + lane;
#pragma unroll 8
          for (int rr = 0; rr < 32; ++rr)
            if (rr < rows_here) __stcs(out + static_cast<int64_t>(rr) * p.C, stg[rr * 33 + lane]);
        }
        __syncwarp();
      }
      if (p.argmax != nullptr && lane < rows_here) p.argmax[wrow0 + lane] = besti;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// returns SIMSEG_ERR_UNSUPPORTED when the shape does not fit the fused kernel (caller falls back to the
// generic inv-norm + GEMM + argmax sequence — still CUDA, still this library)
int patch_sim_fused_impl(Ctx* ctx, const void* patches, int64_t rows, int E, const void* text, int C, int normalize,
                         float* sim, int32_t* argmax, cudaStream_t st) {
  if (E % 64 != 0 || C > 256 || C < 1) return SIMSEG_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(patches) & 15) || (reinterpret_cast<uintptr_t>(text) & 15)) return SIMSEG_ERR_UNSUPPORTED;
  PatchSimParams p{};
  p.rows = rows; p.C = C; p.E = E; p.normalize = normalize;
  p.npad = ((C + 15) / 16) * 16;
  p.kblocks = E / 64;
  p.tiles = static_cast<int>(cdiv(rows, kPsBM));
  p.acc_stride = static_cast<uint32_t>(((p.npad + 31) / 32) * 32);
  uint32_t cols = 32;
  while (cols < 2 * p.acc_stride) cols <<= 1;
  p.tmem_cols = cols;
  p.sim = sim; p.argmax = argmax;
  const int kMaxSmem = 232448;
  const int fixed = 1024 + p.kblocks * p.npad * 128 + kPsStagingBytes + 2 * kPsBM * 4 + 256;
  int stages = (kMaxSmem - fixed) / kPsStageBytes;
  if (stages > kPsMaxStages) stages = kPsMaxStages;
  if (stages < 2) return SIMSEG_ERR_UNSUPPORTED;
  p.stages = stages;
  const int smem_bytes = fixed + stages * kPsStageBytes;
  static int max_set = 0;
  if (smem_bytes > max_set) {
    SIMSEG_CUDA(cudaFuncSetAttribute(patch_sim_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    max_set = smem_bytes;
  }
  CUtensorMap tp, tt;
  int rc = make_tmap(&tp, patches, 2, rows, E, E, 64, kPsBM);
  if (rc) return rc;
  rc = make_tmap(&tt, text, 2, C, E, E, 64, p.npad);
  if (rc) return rc;
  const int grid = p.tiles < ctx->num_sms ? p.tiles : ctx->num_sms;
  patch_sim_kernel<<<grid, kPsThreads, smem_bytes, st>>>(tp, tt, p);
  ctx->launches++;
  SIMSEG_LAUNCH_CHECK();
  return SIMSEG_OK;
}

}  // namespace simseg
