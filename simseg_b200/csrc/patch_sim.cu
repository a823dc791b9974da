// Fused dense patch-text similarity map (tools/seg_evaluation.py:111-112,136-139 of the reference), bf16 inputs:
//
//   sim[r, c] = (patch[r, :] / max(||patch[r, :]||, 1e-12)) . text[c, :]        argmax[r] = first argmax_c sim[r, c]
//
// ONE pass over HBM: every patch row is read once (TMA -> 128B-swizzled smem ring) and every similarity is written
// once.  The class-text matrix ([C, E] bf16, C <= 256) is loaded once per CTA and stays resident in shared memory as
// the B operand; the contraction runs on tcgen05 (128 x Cpad x 16 MMAs, fp32 accumulators in TMEM, two accumulator
// stages) while four "norm" warps read the same smem stages to accumulate the row sums of squares, so the row
// normalisation costs no extra HBM traffic.  The epilogue scales by 1/||patch||, takes the row argmax and writes the map:
//   * C % 4 != 0 (171, 81, 21, 150 classes ...): rows of sim are not 16-byte aligned, but any 16 consecutive rows are one
//     contiguous, 16-byte aligned range of 64*C bytes.  Each lane-quarter warp lays its rows out in shared memory exactly
//     as they lie in HBM (row stride C words: conflict-free for odd C and C = 2 mod 4) and one bulk copy
//     (cp.async.bulk.global.shared::cta) writes the range — no transposition, no per-row address arithmetic, full lines.
//     Two passes of 16 rows per tile keep the staging at 64*C bytes per lane quarter; the second pass re-reads TMEM.
//     Selected for large maps with many classes (CTA-pair mode), where the transposing epilogue is the bottleneck.
//   * otherwise: 32x32 blocks are transposed through padded smem and stored as 128-byte row segments.
//
// Persistent CTAs (one per SM, 14 warps):  warp 0 TMA producer | warp 1 MMA issuer | warps 2-5 row norms |
// warps 6-13 epilogue (two per TMEM lane quarter, taking alternate 32-column chunks; the row argmax of the pair is
// combined through shared memory).
// Algorithmic HBM bytes per row: E*2 + C*4 + 4  (DESIGN.md).
#include "common.cuh"
#include "sm100.cuh"

#include <cstdlib>

namespace simseg {

using namespace sm100;

constexpr int kPsBM = 128;
constexpr int kPsStageBytes = kPsBM * 128;          // 128 rows x 64 bf16
constexpr int kPsThreads = 448;            // 14 warps: TMA, MMA, 4 row-norm, 8 epilogue
constexpr int kPsStagingBytes = 8 * 32 * 33 * 4;   // one padded 32x32 fp32 transpose tile per epilogue warp
constexpr int kPsMaxStages = 8;

struct PatchSimParams {
  int64_t rows;
  int32_t C, E, npad, kblocks, stages, tiles, normalize;
  int32_t bulk;            // 1: contiguous bulk-copy epilogue (C % 4 != 0), 0: transposing epilogue
  int32_t staging_bytes;   // epilogue staging area (multiple of 128)
  uint32_t tmem_cols, acc_stride;
  float* sim;
  int32_t* argmax;
};

// CTAS == 2: a CTA pair works on 256 patch rows; each CTA stages its own 128 rows and HALF of the class-text matrix
// (so C = 171 leaves room for a 7-stage ring instead of 2), the leader issues 256 x Cpad x 16 MMAs (cta_group::2).
// Each CTA's TMA loads signal its OWN barriers (its norm warps read the stages locally); the peer's otherwise idle
// warp 1 relays "stage landed" to the leader, whose MMA thread waits for both halves.
template <int CTAS>
__global__ void __launch_bounds__(kPsThreads, 1)
patch_sim_kernel(const __grid_constant__ CUtensorMap tmap_p, const __grid_constant__ CUtensorMap tmap_t, const PatchSimParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
  const int text_rows = p.npad / CTAS;                                // class rows staged by this CTA
  const int text_kb_bytes = text_rows * 128;
  uint8_t* s_text = smem;
  uint8_t* s_ring = s_text + p.kblocks * text_kb_bytes;
  float* s_stage = reinterpret_cast<float*>(s_ring + p.stages * kPsStageBytes);
  float* s_inv = s_stage + p.staging_bytes / 4;                       // [2][128]
  float* s_arg = s_inv + 2 * kPsBM;                                  // [4 quarters][32 rows][2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_arg + 4 * 64);
  uint64_t* full_bar = bars;                                         // [stages]  TMA -> MMA + norm (local data)
  uint64_t* empty_bar = full_bar + kPsMaxStages;                     // [stages]  MMA commit + 4 norm warps -> TMA
  uint64_t* acc_full = empty_bar + kPsMaxStages;                     // [2] MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;                                // [2] epilogue -> norm (and MMA when CTAS == 1)
  uint64_t* norm_full = acc_empty + 2;                               // [2] norm -> epilogue
  uint64_t* text_bar = norm_full + 2;
  uint64_t* peer_full = text_bar + 1;                                // [stages] leader only: the peer's stage landed
  uint64_t* peer_text = peer_full + kPsMaxStages;                    //          leader only: the peer's text landed
  uint64_t* pair_empty = peer_text + 1;                              // [2] leader only: both CTAs' epilogues done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pair_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = (CTAS == 2) ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const int first_tile = static_cast<int>(blockIdx.x) / CTAS;
  const int tile_stride = static_cast<int>(gridDim.x) / CTAS;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_p);
    prefetch_tmap(&tmap_t);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 5);
      mbar_init(&peer_full[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 8);
      mbar_init(&norm_full[s], 4);
      mbar_init(&pair_empty[s], 16);
    }
    mbar_init(text_bar, 1);
    mbar_init(peer_text, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CTAS == 2) { tmem_alloc_2sm(tmem_slot, p.tmem_cols); tmem_relinquish_2sm(); }
    else { tmem_alloc(tmem_slot, p.tmem_cols); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =============================== TMA producer (warp-uniform loop, elected lane issues) ===============================
    if (elect_one()) {
      mbar_arrive_expect_tx(text_bar, static_cast<uint32_t>(p.kblocks * text_kb_bytes));
      for (int kb = 0; kb < p.kblocks; ++kb)
        tma_load_2d_hint(s_text + kb * text_kb_bytes, &tmap_t, text_bar, kb * 64, static_cast<int>(rank) * text_rows, kEvictLast);
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = first_tile; tile < p.tiles; tile += tile_stride) {
      const int m0 = tile * (kPsBM * CTAS) + static_cast<int>(rank) * kPsBM;
      for (int kb = 0; kb < p.kblocks; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_bar[stage], kPsStageBytes);
          tma_load_2d_hint(s_ring + stage * kPsStageBytes, &tmap_p, &full_bar[stage], kb * 64, m0, kEvictFirst);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // =============================== MMA issuer ===============================
      const uint32_t idesc = make_idesc(1u, 0u, 0u, kPsBM * CTAS, static_cast<uint32_t>(p.npad));
      mbar_wait(text_bar, 0);
      if (CTAS == 2) mbar_wait(peer_text, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint64_t adesc0 = make_smem_desc_sw128(smem_u32(s_ring), 16, 1024);
      const uint64_t bdesc0 = make_smem_desc_sw128(smem_u32(s_text), 16, 1024);
      for (int tile = first_tile; tile < p.tiles; tile += tile_stride) {
        if (CTAS == 2) mbar_wait(&pair_empty[acc], acc_phase ^ 1); else mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * p.acc_stride;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          if (CTAS == 2) mbar_wait(&peer_full[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t ad = adesc0 + static_cast<uint64_t>(stage * (kPsStageBytes >> 4));
            const uint64_t bd = bdesc0 + static_cast<uint64_t>(kb * (text_kb_bytes >> 4));
            const uint32_t first = kb > 0 ? 1u : 0u;
            if (CTAS == 2) {
              umma_f16_2sm(d_tmem, ad, bd, idesc, first);
              umma_f16_2sm(d_tmem, ad + 2, bd + 2, idesc, 1u);
              umma_f16_2sm(d_tmem, ad + 4, bd + 4, idesc, 1u);
              umma_f16_2sm(d_tmem, ad + 6, bd + 6, idesc, 1u);
              umma_commit_2sm(&empty_bar[stage]);
              if (kb == p.kblocks - 1) umma_commit_2sm(&acc_full[acc]);
            } else {
              umma_f16(d_tmem, ad, bd, idesc, first);
              umma_f16(d_tmem, ad + 2, bd + 2, idesc, 1u);
              umma_f16(d_tmem, ad + 4, bd + 4, idesc, 1u);
              umma_f16(d_tmem, ad + 6, bd + 6, idesc, 1u);
              umma_commit(&empty_bar[stage]);
              if (kb == p.kblocks - 1) umma_commit(&acc_full[acc]);
            }
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    } else {
      // =============================== peer: relay "landed" to the leader's MMA thread ===============================
      const uint32_t r_text = mapa_shared(smem_u32(peer_text), 0);
      const uint32_t r_full = mapa_shared(smem_u32(&peer_full[0]), 0);
      mbar_wait(text_bar, 0);
      if (lane == 0) mbar_arrive_cluster(r_text);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = first_tile; tile < p.tiles; tile += tile_stride) {
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          if (lane == 0) mbar_arrive_cluster(r_full + stage * 8);
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
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
    for (int tile = first_tile; tile < p.tiles; tile += tile_stride) {
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
  } else if (p.bulk) {
    // =============================== epilogue, contiguous bulk stores ===============================
    // The two warps of a TMEM lane quarter take alternate 32-column chunks and fill ONE [16 rows][C] staging block laid
    // out as in HBM (two passes per tile: lanes 0-15 stage their rows in pass 0, lanes 16-31 in pass 1, which re-reads
    // TMEM); warp `half == 0` then issues the bulk copy.  Named barrier 1+quarter (64 threads) orders the two warps.
    // A bulk copy takes ~1.5 us to finish reading its block (measured: four passes of 8 rows through two alternating
    // blocks ran 30 % slower), so the wait is paid twice per tile and hidden behind the next tile's MMAs; that needs
    // many tiles per CTA, which is why the host selects this path for large maps only.
    const int quarter = warp & 3;                             // TMEM lane quarter (rows quarter*32 .. +32 of the tile)
    const int half = (warp - 6) >> 2;                         // 0: even 32-column chunks (and the stores), 1: odd chunks
    float* stg = s_stage + quarter * (16 * p.C);              // [16 rows][C] fp32
    float* my_row = stg + (lane & 15) * p.C;
    float* s_best = s_arg + quarter * 64;                     // [32 rows][value, index bits] handed from half 1 to half 0
    const int row_in_tile = quarter * 32 + lane;
    const int nchunks = (p.C + 31) >> 5;
    const uint32_t r_pair_empty = (CTAS == 2) ? mapa_shared(smem_u32(&pair_empty[0]), 0) : 0u;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = first_tile; tile < p.tiles; tile += tile_stride) {
      const int64_t m0 = static_cast<int64_t>(tile) * (kPsBM * CTAS) + static_cast<int64_t>(rank) * kPsBM;
      const int64_t wrow0 = m0 + quarter * 32;                // first global row of this quarter
      mbar_wait(&norm_full[acc], acc_phase);
      const float inv = s_inv[acc * kPsBM + row_in_tile];
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + acc * p.acc_stride + (static_cast<uint32_t>(quarter * 32) << 16);
      float best = -INFINITY;
      int besti = 0;
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {
        const bool mine = (lane >> 4) == pass;
        if (half == 0 && lane == 0) tma_store_wait_read<0>(); // the previous bulk copy has finished reading the block
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
        bool released = pass == 0;
        for (int c = half; c < nchunks; c += 2) {
          const int c0 = c * 32;
          uint32_t r[32];
          tmem_ld_32x32(t_row + c0, r);
          tmem_ld_wait();
          if (pass == 1 && c + 2 >= nchunks) {
            // the last TMEM read of this warp is in registers: hand the accumulator back before the stores
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              mbar_arrive(&acc_empty[acc]);
              if (CTAS == 2) mbar_arrive_cluster(r_pair_empty + acc * 8);
            }
            released = true;
          }
          if (c0 + 32 <= p.C) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float v = __uint_as_float(r[j]) * inv;
              if (pass == 0 && v > best) { best = v; besti = c0 + j; }
              if (mine) my_row[c0 + j] = v;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float v = __uint_as_float(r[j]) * inv;
              if (c0 + j < p.C) {
                if (pass == 0 && v > best) { best = v; besti = c0 + j; }
                if (mine) my_row[c0 + j] = v;
              }
            }
          }
        }
        if (!released) {                                      // this warp had no chunk (C <= 32 and half == 1)
          if (lane == 0) {
            mbar_arrive(&acc_empty[acc]);
            if (CTAS == 2) mbar_arrive_cluster(r_pair_empty + acc * 8);
          }
        }
        if (pass == 1 && half == 1 && p.argmax != nullptr) {
          s_best[2 * lane] = best;
          s_best[2 * lane + 1] = __int_as_float(besti);
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");      // both warps' columns of this pass are staged
        if (half == 0) {
          const int64_t grow = wrow0 + pass * 16;
          const int64_t left = p.rows - grow;
          const int valid = left < 16 ? static_cast<int>(left) : 16;         // may be <= 0 in the ragged last tile
          if (valid > 0) {
            float* out = p.sim + grow * p.C;
            const int n = valid * p.C;
            if ((n & 3) == 0) {
              if (lane == 0) {
                bulk_store_1d(out, stg, static_cast<uint32_t>(n) * 4u);
                tma_store_commit();
              }
            } else {                                          // ragged tail whose byte count is not a multiple of 16
              for (int i = lane; i < n; i += 32) __stcs(out + i, stg[i]);
              __syncwarp();
            }
          }
        }
      }
      // first-max argmax over both halves of the row (half 1 handed its candidate over before the last barrier)
      if (half == 0 && p.argmax != nullptr) {
        const float ob = s_best[2 * lane];
        const int oi = __float_as_int(s_best[2 * lane + 1]);
        if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
        if (lane < p.rows - wrow0) p.argmax[wrow0 + lane] = besti;
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (half == 0 && lane == 0) tma_store_wait<0>();
  } else {
    // =============================== epilogue, transposing (C % 4 == 0) ===============================
    const int quarter = warp & 3;                             // TMEM lane quarter (rows quarter*32 .. +32 of the tile)
    const int half = (warp - 6) >> 2;                         // 0: even 32-column chunks, 1: odd chunks
    float* stg = s_stage + (warp - 6) * (32 * 33);
    float* s_best = s_arg + quarter * 64;                     // [32 rows][value, index bits] handed from half 1 to half 0
    const int row_in_tile = quarter * 32 + lane;
    const int nchunks = (p.npad + 31) >> 5;
    const uint32_t r_pair_empty = (CTAS == 2) ? mapa_shared(smem_u32(&pair_empty[0]), 0) : 0u;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = first_tile; tile < p.tiles; tile += tile_stride) {
      const int64_t m0 = static_cast<int64_t>(tile) * (kPsBM * CTAS) + static_cast<int64_t>(rank) * kPsBM;
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
      bool released = false;
      for (int c = half; c < nchunks; c += 2) {
        const int c0 = c * 32;
        uint32_t r[32];
        tmem_ld_32x32(t_row + c0, r);
        tmem_ld_wait();
        if (c + 2 >= nchunks) {
          // all TMEM reads of this warp are in registers: release the accumulator before the stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&acc_empty[acc]);
            if (CTAS == 2) mbar_arrive_cluster(r_pair_empty + acc * 8);
          }
          released = true;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float v = __uint_as_float(r[j]) * inv;
          if (c0 + j < p.C && v > best) { best = v; besti = c0 + j; }
          stg[lane * 33 + j] = v;
        }
        __syncwarp();
        if (c0 + lane < p.C) {
          float* out = p.sim + wrow0 * p.C + c0 + lane;
          const float* src = stg + lane;
          if (rows_here == 32) {
#pragma unroll
            for (int rr = 0; rr < 32; ++rr) __stcs(out + rr * p.C, src[rr * 33]);
          } else {
            for (int rr = 0; rr < rows_here; ++rr) __stcs(out + rr * p.C, src[rr * 33]);
          }
        }
        __syncwarp();
      }
      if (!released) {                                        // this warp had no chunk (C <= 32 and half == 1)
        if (lane == 0) {
          mbar_arrive(&acc_empty[acc]);
          if (CTAS == 2) mbar_arrive_cluster(r_pair_empty + acc * 8);
        }
      }
      // first-max argmax over both halves of the row: half 1 hands (value, index) to half 0 of the same lane quarter
      if (p.argmax != nullptr) {
        if (half == 1) {
          s_best[2 * lane] = best;
          s_best[2 * lane + 1] = __int_as_float(besti);
        }
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
        if (half == 0) {
          const float ob = s_best[2 * lane];
          const int oi = __float_as_int(s_best[2 * lane + 1]);
          if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
          if (lane < rows_here) p.argmax[wrow0 + lane] = besti;
        }
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");     // s_best may be rewritten only after it was read
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    if (CTAS == 2) tmem_dealloc_2sm(tmem_base, p.tmem_cols); else tmem_dealloc(tmem_base, p.tmem_cols);
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
  p.acc_stride = static_cast<uint32_t>(((p.npad + 31) / 32) * 32);
  uint32_t cols = 32;
  while (cols < 2 * p.acc_stride) cols <<= 1;
  p.tmem_cols = cols;
  p.sim = sim; p.argmax = argmax;
  const int kMaxSmem = 232448;
  // contiguous bulk-copy epilogue: rows of the map not 16-byte aligned, many classes (CTA-pair mode, where the transposing
  // epilogue limits the kernel) and enough tiles per CTA pair to hide the bulk copies' latency behind the next tiles
  p.bulk = 0;
  p.staging_bytes = kPsStagingBytes;
  auto stages_for = [&](int ctas) {
    const int fixed = 1024 + p.kblocks * (p.npad / ctas) * 128 + p.staging_bytes + 2 * kPsBM * 4 + 1024 + 512;
    int stages = (kMaxSmem - fixed) / kPsStageBytes;
    return stages > kPsMaxStages ? kPsMaxStages : stages;
  };
  // a CTA pair splits the text matrix: used when a single CTA could not keep >= 5 patch stages (80 KB) in flight
  const char* force = getenv("SIMSEG_PATCH_SIM_CTAS");
  int ctas = (stages_for(1) >= 5 || rows <= kPsBM) ? 1 : 2;
  if (force) ctas = atoi(force) == 2 ? 2 : 1;
  const char* force_bulk = getenv("SIMSEG_PATCH_SIM_BULK");            // 0 / 1 override the heuristic (tests, A/B timing)
  bool want_bulk = ctas == 2 && cdiv(rows, kPsBM * 2) >= 8 * (ctx->num_sms / 2);
  if (force_bulk) want_bulk = atoi(force_bulk) != 0;
  if (want_bulk && C % 4 != 0 && (reinterpret_cast<uintptr_t>(sim) & 15) == 0) {
    p.bulk = 1;
    p.staging_bytes = ((4 * 16 * C * 4 + 127) / 128) * 128;
    if (stages_for(ctas) < 4) {                                        // the row-contiguous staging would starve the patch ring
      p.bulk = 0;
      p.staging_bytes = kPsStagingBytes;
    }
  }
  const int stages = stages_for(ctas);
  if (stages < 2) return SIMSEG_ERR_UNSUPPORTED;
  p.stages = stages;
  p.tiles = static_cast<int>(cdiv(rows, kPsBM * ctas));
  const int smem_bytes = 1024 + p.kblocks * (p.npad / ctas) * 128 + p.staging_bytes + 2 * kPsBM * 4 + 1024 + 512 + stages * kPsStageBytes;
  CUtensorMap tp, tt;
  int rc = make_tmap(&tp, patches, 2, rows, E, E, 64, kPsBM);
  if (rc) return rc;
  rc = make_tmap(&tt, text, 2, C, E, E, 64, p.npad / ctas);
  if (rc) return rc;
  const int units_max = ctx->num_sms / ctas;
  const int units = p.tiles < units_max ? p.tiles : units_max;
  if (ctas == 1) {
    static int max_set = 0;
    if (smem_bytes > max_set) {
      SIMSEG_CUDA(cudaFuncSetAttribute(patch_sim_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
      max_set = smem_bytes;
    }
    patch_sim_kernel<1><<<units, kPsThreads, smem_bytes, st>>>(tp, tt, p);
  } else {
    static int max_set2 = 0;
    if (smem_bytes > max_set2) {
      SIMSEG_CUDA(cudaFuncSetAttribute(patch_sim_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
      max_set2 = smem_bytes;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(units * 2);
    cfg.blockDim = dim3(kPsThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    SIMSEG_CUDA(cudaLaunchKernelEx(&cfg, patch_sim_kernel<2>, tp, tt, p));
  }
  ctx->launches++;
  SIMSEG_LAUNCH_CHECK();
  return SIMSEG_OK;
}

}  // namespace simseg
