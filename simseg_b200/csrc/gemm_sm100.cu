// tcgen05 + TMA GEMM engine for sm_100a.
//
//   D[M,N] = epilogue( sum_k A[m,k] * B[n,k] )     bf16 (kind::f16) or fp32-as-tf32 (kind::tf32) operands,
//                                                  fp32 accumulators in TMEM.
//
// One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1      MMA issuer    (one elected thread, tcgen05.mma cta_group::1, 128 x BN x 16|8 per instruction)
//   warps 2..9  epilogue      (tcgen05.ld 32x32b, fused bias / GELU / residual / dGELU / row-scale / column sums)
// Two TMEM accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.
// Operands may be K-major (x @ W^T forward) or MN-major (dgrad reads W as [K,N]; wgrad reads dy and x as
// [K=tokens, M|N]); only the TMA box pattern and the smem-descriptor strides differ.
// Split-K (wgrad: K = all tokens) accumulates with vector fp32 reductions (red.global.add.v4.f32).
#include "common.cuh"
#include "sm100.cuh"

#include <cstring>
#include <mutex>

namespace simseg {

using namespace sm100;

#ifndef SIMSEG_STAGES256
#define SIMSEG_STAGES256 4
#endif
constexpr int kBM = 128;           // tile rows = TMEM lanes
constexpr int kSwizzleBytes = 128;  // one swizzle atom row = one K block (K-major) / 64 MN elements (bf16)
constexpr int kNumEpiWarps = 8;
constexpr int kStgTileBytes = kBM * 128;   // epilogue staging tile: 128 rows x 64 bf16, SWIZZLE_128B (TMA store/load box)
constexpr int kGemmThreads = 32 * (2 + kNumEpiWarps);

struct GemmParams {
  int64_t M, N, K;
  int32_t a_mn, b_mn;
  int32_t elem_bytes;      // 2 (bf16) or 4 (tf32)
  int32_t m_tiles, n_tiles, splits, kb_total, kb_per_split;
  void* d;
  int64_t ldd;
  int32_t out_bf16;
  int32_t atomic_out;
  const float* bias;
  const void* residual;
  int64_t ld_res;
  int32_t res_bf16;
  void* aux;
  int64_t ld_aux;
  const float* row_scale;
  float* col_sum;
  int32_t vec_ok;          // 16-byte aligned rows for d / residual / aux
  int32_t tma_epi;         // bf16 output through swizzled smem staging + TMA store (coalesced), else direct stores
  void* aux2;              // DGELU only: gelu(aux) written next to the gradient (bf16 [M,N])
  int32_t a_3d, b_3d;      // MN-major operand loaded with one 3-D TMA box per k-block
  int32_t dbg;             // bench-only: 1 = no TMA after the ring is primed, 2 = no MMA (results are garbage)
};

template <int BN>
struct GemmCfg {
  static constexpr int kABytes = kBM * kSwizzleBytes;          // 16 KB
  static constexpr int kBBytes = BN * kSwizzleBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  // ring depth: measured on B200 the BN=256 mainloop is as fast with 3 stages as with 4 (it is L2-fabric bound, not
  // latency bound), which leaves 64 KB for the epilogue staging tiles
  static constexpr int kStages = (BN <= 128) ? 5 : (BN <= 192 ? 4 : 3);
  static constexpr int kAccStride = (BN <= 128) ? 128 : 256;    // TMEM columns between the two accumulators
  static constexpr int kTmemCols = 2 * kAccStride;              // 256 or 512 (power of two)
  static constexpr int kStagingBytes = 4 * kStgTileBytes;       // four [128 rows][128 B] swizzled epilogue tiles
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 /*align slack*/ + 256 /*barriers*/ + BN * 4;
};


// ------------------------------------------------------------------------------------------------
// Coalesced epilogue for bf16 outputs: TMEM -> registers -> fused math -> 128B-swizzled smem tile [128 rows][64 cols]
// -> TMA store.  All 8 epilogue warps work on one 64-column chunk at a time (warp = TMEM lane quarter x 32-column
// half); thread = one output row, so every smem access is a conflict-free 16-byte chunk of the thread's own row.
//   EPI_NONE      out = acc (+bias)                                   staging tiles rotate over 4 slots
//   EPI_BIAS_GELU aux = bf16(acc+bias) ; out = gelu(aux)              two stores per chunk, 2 x 2 slots
//   EPI_DGELU     out = acc * gelu'(aux) ; aux2 = gelu(aux) ; column sums of out
//                 aux tiles are TMA-LOADED two chunks ahead into a 3-slot ring; gelu(aux) overwrites the tile in place
//                 and is TMA-stored from there (the backward pass never runs a separate GELU recompute kernel)
__device__ __forceinline__ void epi_bar(int id) { asm volatile("bar.sync %0, 256;" ::"r"(id) : "memory"); }

template <int BN, int EPI>
__device__ __forceinline__ void epilogue_tma(const GemmParams& p, const CUtensorMap& tmap_d, const CUtensorMap& tmap_x,
                                             const CUtensorMap& tmap_x2, uint8_t* stg, uint64_t* acc_full, uint64_t* acc_empty,
                                             uint64_t* aux_full, float* s_cs, uint32_t tmem_base, int total_tiles, int tiles_mn,
                                             int warp, int lane) {
  using Cfg = GemmCfg<BN>;
  constexpr int kCh = BN / 64;
  const int ew = warp - 2;
  const int quarter = warp & 3;
  const int half = ew >> 2;
  const int r = quarter * 32 + lane;
  const int sw = r & 7;
  const bool elected = (ew == 0 && lane == 0);
  const uint32_t row_off = static_cast<uint32_t>(r) * 128;
  const int tid = ew * 32 + lane;                                   // 0..255 among the epilogue threads
  uint32_t cc = 0;                                                  // flat chunk counter over (tile, chunk)

  auto chunk_coords = [&](uint32_t f, int& m0, int& ncol) -> bool {
    const int tile = static_cast<int>(blockIdx.x) + static_cast<int>(f / kCh) * static_cast<int>(gridDim.x);
    if (tile >= total_tiles) return false;
    const int mn = tile % tiles_mn;
    m0 = (mn / p.n_tiles) * kBM;
    ncol = (mn % p.n_tiles) * BN + static_cast<int>(f % kCh) * 64;
    return true;
  };
  if (EPI == SIMSEG_EPI_DGELU && elected) {
    for (uint32_t f = 0; f < 2; ++f) {
      int m0, nc;
      if (chunk_coords(f, m0, nc)) {
        mbar_arrive_expect_tx(&aux_full[f % 3], kStgTileBytes);
        tma_load_2d(stg + (1 + f % 3) * kStgTileBytes, &tmap_x, &aux_full[f % 3], nc, m0);
      }
    }
  }

  int acc = 0;
  uint32_t acc_phase = 0;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int mn = tile % tiles_mn;
    const int m0 = (mn / p.n_tiles) * kBM;
    const int n0 = (mn % p.n_tiles) * BN;
    mbar_wait(&acc_full[acc], acc_phase);
    tc_fence_after();
    const uint32_t t_row = tmem_base + acc * Cfg::kAccStride + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < kCh; ++c, ++cc) {
      const int col_in_tile = c * 64 + half * 32;
      const int col0 = n0 + col_in_tile;
      uint32_t rr[32];
      tmem_ld_32x32(t_row + col_in_tile, rr);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rr[j]);
      if (p.bias != nullptr) {
        if (col0 + 32 <= p.N) {
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = __ldg(b4 + j);
            v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) if (col0 + j < p.N) v[j] += __ldg(p.bias + col0 + j);
        }
      }
      int out_slot, x_slot;
      if (EPI == SIMSEG_EPI_BIAS_GELU) { x_slot = 2 * (cc & 1); out_slot = x_slot + 1; }
      else if (EPI == SIMSEG_EPI_DGELU) { out_slot = 0; x_slot = 1 + static_cast<int>(cc % 3); }
      else { out_slot = static_cast<int>(cc & 3); x_slot = 0; }
      uint8_t* so = stg + out_slot * kStgTileBytes + row_off;
      uint8_t* sx = stg + x_slot * kStgTileBytes + row_off;
      uint32_t xo[16];                                              // second bf16 output of this row chunk (aux / aux2)
      if (EPI == SIMSEG_EPI_BIAS_GELU) {
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          xo[j >> 1] = pack_bf16(v[j], v[j + 1]);                   // pre-activation, bf16 — exactly what backward reads
          v[j] = gelu_erf(bf16_lo(xo[j >> 1]));
          v[j + 1] = gelu_erf(bf16_hi(xo[j >> 1]));
        }
      } else if (EPI == SIMSEG_EPI_DGELU) {
        mbar_wait(&aux_full[cc % 3], (cc / 3) & 1);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const uint4 u = *reinterpret_cast<const uint4*>(sx + (((half * 4 + q4) ^ sw) << 4));
          const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float g0, a0, g1, a1;
            gelu_erf_both(bf16_lo(w[e]), a0, g0);
            gelu_erf_both(bf16_hi(w[e]), a1, g1);
            v[8 * q4 + 2 * e] *= g0;
            v[8 * q4 + 2 * e + 1] *= g1;
            xo[4 * q4 + e] = pack_bf16(a0, a1);
          }
        }
      }
      uint32_t o[16];
#pragma unroll
      for (int j = 0; j < 32; j += 2) o[j >> 1] = pack_bf16(v[j], v[j + 1]);
      if (EPI == SIMSEG_EPI_DGELU && p.col_sum != nullptr) {
        // column sums over this warp's 32 rows (butterfly transpose-reduce), then one shared-memory add per column
#pragma unroll
        for (int o2 = 16; o2 >= 1; o2 >>= 1) {
          const bool upper = (lane & o2) != 0;
#pragma unroll
          for (int j = 0; j < o2; ++j) {
            const float mine = upper ? v[j + o2] : v[j];
            const float send = upper ? v[j] : v[j + o2];
            v[j] = mine + __shfl_xor_sync(0xffffffffu, send, o2);
          }
        }
        atomicAdd(&s_cs[col_in_tile + lane], v[0]);
      }
      // ---- staging slot must have been read by the TMA store that last used it
      if (elected) {
        if (EPI == SIMSEG_EPI_BIAS_GELU) tma_store_wait_read<1>();
        else if (EPI == SIMSEG_EPI_DGELU) tma_store_wait_read<0>();
        else tma_store_wait_read<3>();
        if (EPI == SIMSEG_EPI_DGELU) {
          int m2, nc2;
          if (chunk_coords(cc + 2, m2, nc2)) {                       // slot (cc+2)%3 == (cc-1)%3: just released
            mbar_arrive_expect_tx(&aux_full[(cc + 2) % 3], kStgTileBytes);
            tma_load_2d(stg + (1 + (cc + 2) % 3) * kStgTileBytes, &tmap_x, &aux_full[(cc + 2) % 3], nc2, m2);
          }
        }
      }
      epi_bar(1);
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const uint32_t off = ((half * 4 + q4) ^ sw) << 4;
        *reinterpret_cast<uint4*>(so + off) = make_uint4(o[4 * q4], o[4 * q4 + 1], o[4 * q4 + 2], o[4 * q4 + 3]);
        if ((EPI == SIMSEG_EPI_BIAS_GELU && p.aux != nullptr) || (EPI == SIMSEG_EPI_DGELU && p.aux2 != nullptr))
          *reinterpret_cast<uint4*>(sx + off) = make_uint4(xo[4 * q4], xo[4 * q4 + 1], xo[4 * q4 + 2], xo[4 * q4 + 3]);
      }
      fence_proxy_async_smem();
      epi_bar(2);
      if (elected) {
        tma_store_2d(&tmap_d, stg + out_slot * kStgTileBytes, n0 + c * 64, m0);
        if (EPI == SIMSEG_EPI_BIAS_GELU && p.aux != nullptr) tma_store_2d(&tmap_x, stg + x_slot * kStgTileBytes, n0 + c * 64, m0);
        if (EPI == SIMSEG_EPI_DGELU && p.aux2 != nullptr) tma_store_2d(&tmap_x2, stg + x_slot * kStgTileBytes, n0 + c * 64, m0);
        tma_store_commit();
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&acc_empty[acc]);
    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    if (EPI == SIMSEG_EPI_DGELU && p.col_sum != nullptr) {
      // every warp's shared-memory adds of this tile happened before the last epi_bar(1): flush and clear
      if (tid < BN) {
        const float x = s_cs[tid];
        s_cs[tid] = 0.f;
        if (n0 + tid < p.N) atomicAdd(p.col_sum + n0 + tid, x);
      }
      epi_bar(1);
    }
  }
  if (elected) tma_store_wait<0>();
}

// ------------------------------------------------------------------------------------------------
template <int BN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
            const __grid_constant__ CUtensorMap tmap_d, const __grid_constant__ CUtensorMap tmap_x,
            const __grid_constant__ CUtensorMap tmap_x2, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment is required by SWIZZLE_128B (descriptor base_offset = 0)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stg = smem + Cfg::kStages * Cfg::kStageBytes;      // 4 staging tiles (1024-byte aligned)
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg + Cfg::kStagingBytes);
  uint64_t* full_bar = bars;                       // [kStages]  TMA -> MMA
  uint64_t* empty_bar = bars + Cfg::kStages;       // [kStages]  MMA -> TMA
  uint64_t* acc_full = bars + 2 * Cfg::kStages;    // [2]        MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;              // [2]        epilogue -> MMA
  uint64_t* aux_full = acc_empty + 2;              // [3]        TMA (aux-in tiles of the DGELU epilogue) -> epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_full + 3);
  float* s_cs = reinterpret_cast<float*>(bars + 32);           // [BN] per-tile column sums

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], kNumEpiWarps);
    }
    for (int s = 0; s < 3; ++s) mbar_init(&aux_full[s], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  if (threadIdx.x >= 64 && threadIdx.x - 64 < BN) s_cs[threadIdx.x - 64] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = p.m_tiles * p.n_tiles * p.splits;
  const int tiles_mn = p.m_tiles * p.n_tiles;
  const int k_elems = kSwizzleBytes / p.elem_bytes;      // K elements per k-block: 64 (bf16) or 32 (tf32)

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int split = tile / tiles_mn;
        const int mn = tile - split * tiles_mn;
        const int m0 = (mn / p.n_tiles) * kBM;
        const int n0 = (mn % p.n_tiles) * BN;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          if (p.dbg == 1 && (phase != 0 || tile != static_cast<int>(blockIdx.x))) {
            mbar_arrive(&full_bar[stage]);
            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          const int k0 = kb * k_elems;
          if (!p.a_mn) {
            tma_load_2d(sa, &tmap_a, &full_bar[stage], k0, m0);
          } else {
            // MN-major: boxes of (swizzle-row of MN elements) x (k_elems rows... see host: box = {mn_atom, k_rows})
            const int mn_atom = kSwizzleBytes / p.elem_bytes;          // 64 bf16 / 32 tf32 elements
            const int nbox = kBM / mn_atom;
            const int box_bytes = Cfg::kABytes / nbox;
            if (p.a_3d) tma_load_3d(sa, &tmap_a, &full_bar[stage], 0, k0, m0 / mn_atom);
            else
              for (int j = 0; j < nbox; ++j) tma_load_2d(sa + j * box_bytes, &tmap_a, &full_bar[stage], m0 + j * mn_atom, k0);
          }
          if (!p.b_mn) {
            tma_load_2d(sb, &tmap_b, &full_bar[stage], k0, n0);
          } else {
            const int mn_atom = kSwizzleBytes / p.elem_bytes;
            const int nbox = BN / mn_atom;
            const int box_bytes = Cfg::kBBytes / nbox;
            if (p.b_3d) tma_load_3d(sb, &tmap_b, &full_bar[stage], 0, k0, n0 / mn_atom);
            else
              for (int j = 0; j < nbox; ++j) tma_load_2d(sb + j * box_bytes, &tmap_b, &full_bar[stage], n0 + j * mn_atom, k0);
          }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(p.elem_bytes == 2 ? 1u : 2u, p.a_mn, p.b_mn, kBM, BN);
      // K-major: consecutive UMMA_K slices are 32 B apart inside the 128 B swizzle row; 8-row groups 1024 B apart.
      // MN-major: a UMMA_K slice (16 bf16 / 8 tf32 k-rows... = 32 B of K per row-group) spans k-rows of 128 B each,
      //           8-row groups 1024 B apart (SBO), 64-element MN atoms one TMA box apart (LBO).
      const int kk_steps = 4;                                   // 128 B / 32 B
      const uint32_t umma_k = 32 / p.elem_bytes;                // 16 (bf16) / 8 (tf32)
      const uint32_t a_step = p.a_mn ? umma_k * kSwizzleBytes : 32;
      const uint32_t b_step = p.b_mn ? umma_k * kSwizzleBytes : 32;
      const uint32_t k_rows_bytes = (kSwizzleBytes / p.elem_bytes) * kSwizzleBytes;   // one MN-major box: k_elems rows x 128 B
      const uint32_t a_lbo = p.a_mn ? k_rows_bytes : 16;
      const uint32_t b_lbo = p.b_mn ? k_rows_bytes : 16;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int split = tile / tiles_mn;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::kAccStride;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t sb = sa + Cfg::kABytes;
          if (p.dbg == 2 && kb > kb0) {
            mbar_arrive(&empty_bar[stage]);
            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
            continue;
          }
#pragma unroll
          for (int kk = 0; kk < kk_steps; ++kk) {
            const uint64_t adesc = make_smem_desc_sw128(sa + kk * a_step, a_lbo, 1024);
            const uint64_t bdesc = make_smem_desc_sw128(sb + kk * b_step, b_lbo, 1024);
            const uint32_t accum = (kb > kb0 || kk > 0) ? 1u : 0u;
            if (p.elem_bytes == 2) umma_f16(d_tmem, adesc, bdesc, idesc, accum);
            else umma_tf32(d_tmem, adesc, bdesc, idesc, accum);
          }
          umma_commit(&empty_bar[stage]);            // frees this smem slot once the MMAs have read it
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&acc_full[acc]);                 // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // =============================== epilogue ===============================
    if (p.tma_epi) {
      epilogue_tma<BN, EPI>(p, tmap_d, tmap_x, tmap_x2, stg, acc_full, acc_empty, aux_full, s_cs, tmem_base, total_tiles, tiles_mn, warp, lane);
    } else {
    const int ew = warp - 2;                 // 0..7
    const int quarter = warp & 3;            // TMEM lane quarter this warp may access
    const int half = ew >> 2;                // column half handled by this warp
    constexpr int kColsPerWarp = BN / 2;
    constexpr int kChunks = kColsPerWarp / 32;
    static_assert(kColsPerWarp % 32 == 0, "BN must be a multiple of 64");
    const int row_in_tile = quarter * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int split = tile / tiles_mn;
      const int mn = tile - split * tiles_mn;
      const int m0 = (mn / p.n_tiles) * kBM;
      const int n0 = (mn % p.n_tiles) * BN;
      const int64_t row = m0 + row_in_tile;
      const bool row_ok = row < p.M;
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + acc * Cfg::kAccStride + (static_cast<uint32_t>(quarter * 32) << 16);
      float rs = 1.0f;
      if (EPI == SIMSEG_EPI_ROWSCALE) rs = row_ok ? p.row_scale[row] : 0.0f;
#pragma unroll 1
      for (int c = 0; c < kChunks; ++c) {
        const int col_in_tile = half * kColsPerWarp + c * 32;
        const int64_t col0 = n0 + col_in_tile;
        uint32_t r[32];
        tmem_ld_32x32(t_row + col_in_tile, r);
        tmem_ld_wait();
        if (col0 >= p.N) continue;                               // whole chunk is N padding (warp-uniform)
        const bool full_chunk = (col0 + 32 <= p.N);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        // ---- bias
        if (p.bias != nullptr && (split == 0)) {
          if (full_chunk && p.vec_ok) {
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = __ldg(b4 + j);
              v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (col0 + j < p.N) v[j] += __ldg(p.bias + col0 + j);
          }
        }
        // ---- fused math
        if (EPI == SIMSEG_EPI_BIAS_GELU) {
          if (p.aux != nullptr && row_ok) {
            __nv_bfloat16* ap = reinterpret_cast<__nv_bfloat16*>(p.aux) + row * p.ld_aux + col0;
            if (full_chunk && p.vec_ok) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 u;
                u.x = pack_bf16(v[8 * j], v[8 * j + 1]); u.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
                u.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]); u.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
                reinterpret_cast<uint4*>(ap)[j] = u;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (col0 + j < p.N) ap[j] = __float2bfloat16(v[j]);
            }
          }
          // GELU is applied to the bf16-rounded pre-activation, exactly what backward will read back
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_erf(__bfloat162float(__float2bfloat16(v[j])));
        } else if (EPI == SIMSEG_EPI_BIAS_RESIDUAL) {
          if (row_ok) {
            if (p.res_bf16) {
              const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.residual) + row * p.ld_res + col0;
              if (full_chunk && p.vec_ok) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const uint4 u = reinterpret_cast<const uint4*>(rp)[j];
                  v[8 * j] += bf16_lo(u.x); v[8 * j + 1] += bf16_hi(u.x); v[8 * j + 2] += bf16_lo(u.y); v[8 * j + 3] += bf16_hi(u.y);
                  v[8 * j + 4] += bf16_lo(u.z); v[8 * j + 5] += bf16_hi(u.z); v[8 * j + 6] += bf16_lo(u.w); v[8 * j + 7] += bf16_hi(u.w);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) if (col0 + j < p.N) v[j] += __bfloat162float(rp[j]);
              }
            } else {
              const float* rp = reinterpret_cast<const float*>(p.residual) + row * p.ld_res + col0;
              if (full_chunk && p.vec_ok) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 f = reinterpret_cast<const float4*>(rp)[j];
                  v[4 * j] += f.x; v[4 * j + 1] += f.y; v[4 * j + 2] += f.z; v[4 * j + 3] += f.w;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) if (col0 + j < p.N) v[j] += rp[j];
              }
            }
          }
        } else if (EPI == SIMSEG_EPI_DGELU) {
          if (row_ok) {
            const __nv_bfloat16* hp = reinterpret_cast<const __nv_bfloat16*>(p.aux) + row * p.ld_aux + col0;
            if (full_chunk && p.vec_ok) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 u = reinterpret_cast<const uint4*>(hp)[j];
                v[8 * j] *= gelu_erf_grad(bf16_lo(u.x)); v[8 * j + 1] *= gelu_erf_grad(bf16_hi(u.x));
                v[8 * j + 2] *= gelu_erf_grad(bf16_lo(u.y)); v[8 * j + 3] *= gelu_erf_grad(bf16_hi(u.y));
                v[8 * j + 4] *= gelu_erf_grad(bf16_lo(u.z)); v[8 * j + 5] *= gelu_erf_grad(bf16_hi(u.z));
                v[8 * j + 6] *= gelu_erf_grad(bf16_lo(u.w)); v[8 * j + 7] *= gelu_erf_grad(bf16_hi(u.w));
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (col0 + j < p.N) v[j] *= gelu_erf_grad(__bfloat162float(hp[j]));
            }
          }
        } else if (EPI == SIMSEG_EPI_ROWSCALE) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= rs;
        }
        // ---- store
        if (row_ok) {
          if (p.atomic_out) {
            float* dp = reinterpret_cast<float*>(p.d) + row * p.ldd + col0;
            if (full_chunk && p.vec_ok) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dp + 4 * j), "f"(v[4 * j]), "f"(v[4 * j + 1]),
                             "f"(v[4 * j + 2]), "f"(v[4 * j + 3]) : "memory");
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (col0 + j < p.N) atomicAdd(dp + j, v[j]);
            }
          } else if (p.out_bf16) {
            __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(p.d) + row * p.ldd + col0;
            if (full_chunk && p.vec_ok) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 u;
                u.x = pack_bf16(v[8 * j], v[8 * j + 1]); u.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
                u.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]); u.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
                reinterpret_cast<uint4*>(dp)[j] = u;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (col0 + j < p.N) dp[j] = __float2bfloat16(v[j]);
            }
          } else {
            float* dp = reinterpret_cast<float*>(p.d) + row * p.ldd + col0;
            if (full_chunk && p.vec_ok) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                reinterpret_cast<float4*>(dp)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (col0 + j < p.N) dp[j] = v[j];
            }
          }
        }
        // ---- column sums over the 32 rows of this warp (butterfly transpose-reduce, 31 shuffles)
        if (p.col_sum != nullptr) {
          if (!row_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.0f;
          }
          // after the step with offset o, lane keeps the half of its columns selected by (lane & o)
#pragma unroll
          for (int o = 16; o >= 1; o >>= 1) {
            const bool upper = (lane & o) != 0;
#pragma unroll
            for (int j = 0; j < o; ++j) {
              const float mine = upper ? v[j + o] : v[j];
              const float send = upper ? v[j] : v[j + o];
              v[j] = mine + __shfl_xor_sync(0xffffffffu, send, o);
            }
          }
          // lane l now holds the column whose index bits equal l's bits: column = lane
          if (col0 + lane < p.N) atomicAdd(p.col_sum + col0 + lane, v[0]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

// 2-D row-major tensor [rows, cols] (cols contiguous, leading dim ld elements); box = {box_cols, box_rows}
int make_tmap(CUtensorMap* m, const void* ptr, int elem_bytes, int64_t rows, int64_t cols, int64_t ld,
              int box_cols, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled not available (driver entry point lookup failed)"); return SIMSEG_ERR_CUDA; }
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * elem_bytes};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): ptr=%p rows=%lld cols=%lld ld=%lld box=%dx%d", static_cast<int>(r), ptr,
              static_cast<long long>(rows), static_cast<long long>(cols), static_cast<long long>(ld), box_cols, box_rows);
    return SIMSEG_ERR_CUDA;
  }
  return SIMSEG_OK;
}

int make_tmap_mn3d(CUtensorMap* m, const void* ptr, int elem_bytes, int64_t k_rows, int64_t mn_cols, int64_t ld, int box_k,
                   int box_atoms) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled not available (driver entry point lookup failed)"); return SIMSEG_ERR_CUDA; }
  const int atom = kSwizzleBytes / elem_bytes;
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(atom), static_cast<cuuint64_t>(k_rows), static_cast<cuuint64_t>(mn_cols / atom)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * elem_bytes, static_cast<cuuint64_t>(kSwizzleBytes)};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(atom), static_cast<cuuint32_t>(box_k), static_cast<cuuint32_t>(box_atoms)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(3d) failed (%d): ptr=%p k_rows=%lld mn=%lld ld=%lld box_k=%d atoms=%d", static_cast<int>(r),
              ptr, static_cast<long long>(k_rows), static_cast<long long>(mn_cols), static_cast<long long>(ld), box_k, box_atoms);
    return SIMSEG_ERR_CUDA;
  }
  return SIMSEG_OK;
}

template <int BN, int EPI>
static int launch_gemm(Ctx* ctx, const CUtensorMap* tm, const GemmParams& p, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  auto kfn = gemm_kernel<BN, EPI>;
  if (!attr_set) {
    SIMSEG_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  const int total = p.m_tiles * p.n_tiles * p.splits;
  const int grid = total < ctx->num_sms ? total : ctx->num_sms;
  kfn<<<grid, kGemmThreads, Cfg::kSmemBytes, st>>>(tm[0], tm[1], tm[2], tm[3], tm[4], p);
  ctx->launches++;
  SIMSEG_LAUNCH_CHECK();
  return SIMSEG_OK;
}

template <int BN>
static int dispatch_epi(Ctx* ctx, int epi, const CUtensorMap* tm, const GemmParams& p, cudaStream_t st) {
  switch (epi) {
    case SIMSEG_EPI_NONE: return launch_gemm<BN, SIMSEG_EPI_NONE>(ctx, tm, p, st);
    case SIMSEG_EPI_BIAS_GELU: return launch_gemm<BN, SIMSEG_EPI_BIAS_GELU>(ctx, tm, p, st);
    case SIMSEG_EPI_BIAS_RESIDUAL: return launch_gemm<BN, SIMSEG_EPI_BIAS_RESIDUAL>(ctx, tm, p, st);
    case SIMSEG_EPI_DGELU: return launch_gemm<BN, SIMSEG_EPI_DGELU>(ctx, tm, p, st);
    case SIMSEG_EPI_ROWSCALE: return launch_gemm<BN, SIMSEG_EPI_ROWSCALE>(ctx, tm, p, st);
  }
  set_error("unknown epilogue %d", epi);
  return SIMSEG_ERR_INVALID;
}

int gemm_impl(Ctx* ctx, const simseg_gemm_args* a, cudaStream_t st) {
  SIMSEG_CHECK_ARG(a->M > 0 && a->N > 0 && a->K > 0, "gemm: empty problem M=%lld N=%lld K=%lld", (long long)a->M,
                   (long long)a->N, (long long)a->K);
  SIMSEG_CHECK_ARG(a->in_dtype == SIMSEG_BF16 || a->in_dtype == SIMSEG_F32, "gemm: bad in_dtype %d", a->in_dtype);
  const int eb = a->in_dtype == SIMSEG_BF16 ? 2 : 4;
  if (a->in_dtype == SIMSEG_F32 && (a->a_major || a->b_major)) {
    // 32-bit MN-major operands need the 128B_BASE32B shared-memory layout, which this engine does not stage
    set_error("gemm: tf32 operands must be K-major");
    return SIMSEG_ERR_UNSUPPORTED;
  }
  SIMSEG_CHECK_ARG((a->lda * eb) % 16 == 0 && (a->ldb * eb) % 16 == 0, "gemm: lda/ldb rows must be 16-byte multiples");
  SIMSEG_CHECK_ARG((reinterpret_cast<uintptr_t>(a->a) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->b) & 15) == 0,
                   "gemm: A/B must be 16-byte aligned");
  SIMSEG_CHECK_ARG(!(a->accumulate && a->out_dtype != SIMSEG_F32), "gemm: accumulate needs fp32 output");
  if (a->epilogue == SIMSEG_EPI_BIAS_RESIDUAL) SIMSEG_CHECK_ARG(a->residual != nullptr, "gemm: residual missing");
  if (a->epilogue == SIMSEG_EPI_DGELU) SIMSEG_CHECK_ARG(a->aux != nullptr, "gemm: aux (pre-activation) missing");
  if (a->epilogue == SIMSEG_EPI_ROWSCALE) SIMSEG_CHECK_ARG(a->row_scale != nullptr, "gemm: row_scale missing");

  int bn = a->tile_n;
  if (bn == 0) {
    // smallest padding waste first, widest tile on ties (fewer A re-reads, better MMA efficiency)
    const int cand[3] = {256, 192, 128};
    int64_t best = -1;
    for (int c : cand) {
      const int64_t padded = cdiv(a->N, c) * c;
      if (best < 0 || padded < best) { best = padded; bn = c; }
    }
  }
  SIMSEG_CHECK_ARG(bn == 128 || bn == 192 || bn == 256, "gemm: tile_n must be 128/192/256 (got %d)", bn);

  GemmParams p{};
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.a_mn = a->a_major ? 1 : 0; p.b_mn = a->b_major ? 1 : 0;
  p.elem_bytes = eb;
  p.m_tiles = static_cast<int>(cdiv(a->M, kBM));
  p.n_tiles = static_cast<int>(cdiv(a->N, bn));
  const int k_elems = kSwizzleBytes / eb;
  p.kb_total = static_cast<int>(cdiv(a->K, k_elems));
  // split-K only for plain fp32 accumulation outputs with few output tiles (wgrad: K = all tokens).
  // The split count is chosen for wave efficiency: work items (tiles x splits) should fill whole waves of
  // num_sms persistent CTAs; among equally efficient choices the smallest split count wins (fewer atomics).
  int splits = 1;
  const int tiles_mn = p.m_tiles * p.n_tiles;
  const bool can_split = a->epilogue == SIMSEG_EPI_NONE && a->out_dtype == SIMSEG_F32 && a->col_sum == nullptr;
  if (can_split && p.kb_total >= 32) {
    const int max_splits = p.kb_total / 8 < 64 ? p.kb_total / 8 : 64;
    double best_eff = 0.0;
    for (int s = 1; s <= max_splits; ++s) {
      const int kbs = static_cast<int>(cdiv(p.kb_total, s));
      const int s_eff = static_cast<int>(cdiv(p.kb_total, kbs));
      const int64_t items = static_cast<int64_t>(tiles_mn) * s_eff;
      const int64_t waves = cdiv(items, ctx->num_sms);
      // time ~ waves * (k-blocks per item + fixed per-item cost of ~6 k-blocks for prologue/epilogue)
      const double eff = static_cast<double>(tiles_mn) * p.kb_total / (static_cast<double>(waves) * ctx->num_sms * (kbs + 6.0));
      if (eff > best_eff * 1.02) { best_eff = eff; splits = s_eff; }
    }
  }
  p.kb_per_split = static_cast<int>(cdiv(p.kb_total, splits));
  p.splits = static_cast<int>(cdiv(p.kb_total, p.kb_per_split));
  p.d = a->d; p.ldd = a->ldd;
  p.out_bf16 = a->out_dtype == SIMSEG_BF16;
  p.atomic_out = (p.splits > 1 || a->accumulate) ? 1 : 0;
  p.bias = a->bias; p.residual = a->residual; p.ld_res = a->ld_res; p.res_bf16 = a->res_dtype == SIMSEG_BF16;
  p.aux = a->aux; p.ld_aux = a->ld_aux; p.row_scale = a->row_scale; p.col_sum = a->col_sum;
  const int ob = p.out_bf16 ? 2 : 4;
  bool vec = (a->ldd * ob) % 16 == 0 && (reinterpret_cast<uintptr_t>(a->d) & 15) == 0;
  if (a->residual) vec = vec && (a->ld_res * (p.res_bf16 ? 2 : 4)) % 16 == 0 && (reinterpret_cast<uintptr_t>(a->residual) & 15) == 0;
  if (a->aux) vec = vec && (a->ld_aux * 2) % 16 == 0 && (reinterpret_cast<uintptr_t>(a->aux) & 15) == 0;
  if (a->bias) vec = vec && (reinterpret_cast<uintptr_t>(a->bias) & 15) == 0;
  p.vec_ok = vec ? 1 : 0;
  p.dbg = a->reserved & 3;

  if (p.splits > 1 && !a->accumulate) {
    // split-K partial sums are reduced with atomics: start from zero
    SIMSEG_CUDA(cudaMemset2DAsync(a->d, a->ldd * 4, 0, a->N * 4, a->M, st));
  }

  CUtensorMap tm[5];
  memset(tm, 0, sizeof(tm));
  CUtensorMap& ta = tm[0];
  CUtensorMap& tb = tm[1];
  const int mn_atom = kSwizzleBytes / eb;
  int rc;
  // coalesced TMA-store epilogue: bf16 output, no split-K / accumulation, 16-byte aligned rows
  const bool epi_ok = a->epilogue == SIMSEG_EPI_NONE || a->epilogue == SIMSEG_EPI_BIAS_GELU || a->epilogue == SIMSEG_EPI_DGELU;
  auto row_ok16 = [](const void* ptr, int64_t ld) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld * 2) % 16 == 0; };
  bool tma_epi = epi_ok && p.out_bf16 && !p.atomic_out && row_ok16(a->d, a->ldd) && (a->reserved & 8) == 0;
  if (a->aux) tma_epi = tma_epi && row_ok16(a->aux, a->ld_aux);
  if (a->aux2) tma_epi = tma_epi && row_ok16(a->aux2, a->ld_aux2) && a->epilogue == SIMSEG_EPI_DGELU;
  if (a->bias) tma_epi = tma_epi && (reinterpret_cast<uintptr_t>(a->bias) & 15) == 0;
  if (a->epilogue != SIMSEG_EPI_DGELU && a->col_sum) tma_epi = false;
  SIMSEG_CHECK_ARG(!(a->aux2 && !tma_epi), "gemm: aux2 needs the bf16 TMA epilogue (DGELU, bf16 out, 16-byte aligned rows)");
  p.tma_epi = tma_epi ? 1 : 0;
  p.aux2 = a->aux2;
  if (tma_epi) {
    if ((rc = make_tmap(&tm[2], a->d, 2, a->M, a->N, a->ldd, 64, kBM))) return rc;
    if (a->aux && (rc = make_tmap(&tm[3], a->aux, 2, a->M, a->N, a->ld_aux, 64, kBM))) return rc;
    if (a->aux2 && (rc = make_tmap(&tm[4], a->aux2, 2, a->M, a->N, a->ld_aux2, 64, kBM))) return rc;
  }
  // MN-major operands whose MN extent is a whole number of 128-byte atoms take ONE 3-D box per k-block
  // (atoms past the matrix edge are zero-filled by TMA); ragged extents keep one 2-D box per atom.
  const bool no3d = (a->reserved & 4) != 0;
  p.a_3d = (p.a_mn && a->M % mn_atom == 0 && !no3d) ? 1 : 0;
  p.b_3d = (p.b_mn && a->N % mn_atom == 0 && !no3d) ? 1 : 0;
  if (!p.a_mn) rc = make_tmap(&ta, a->a, eb, a->M, a->K, a->lda, k_elems, kBM);
  else if (p.a_3d) rc = make_tmap_mn3d(&ta, a->a, eb, a->K, a->M, a->lda, k_elems, kBM / mn_atom);
  else rc = make_tmap(&ta, a->a, eb, a->K, a->M, a->lda, mn_atom, k_elems);
  if (rc) return rc;
  if (!p.b_mn) rc = make_tmap(&tb, a->b, eb, a->N, a->K, a->ldb, k_elems, bn);
  else if (p.b_3d) rc = make_tmap_mn3d(&tb, a->b, eb, a->K, a->N, a->ldb, k_elems, bn / mn_atom);
  else rc = make_tmap(&tb, a->b, eb, a->K, a->N, a->ldb, mn_atom, k_elems);
  if (rc) return rc;

  switch (bn) {
    case 128: return dispatch_epi<128>(ctx, a->epilogue, tm, p, st);
    case 192: return dispatch_epi<192>(ctx, a->epilogue, tm, p, st);
    default: return dispatch_epi<256>(ctx, a->epilogue, tm, p, st);
  }
}

}  // namespace simseg
