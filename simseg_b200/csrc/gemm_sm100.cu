// tcgen05 + TMA GEMM engine for sm_100a.
//
//   D[M,N] = epilogue( sum_k A[m,k] * B[n,k] )     bf16 (kind::f16) or fp32-as-tf32 (kind::tf32) operands,
//                                                  fp32 accumulators in TMEM.
//
// Persistent, warp-specialised, one CTA per SM; optionally two CTAs (the two SMs of a TPC) form a PAIR that works on
// one 256 x BN tile with tcgen05 cta_group::2: each CTA stages its own 128 rows of A and HALF of the B columns, the
// leader issues 256 x BN x 16 MMAs that read both CTAs' shared memory and write 128 accumulator lanes into each
// CTA's TMEM.  Per flop that is 1/3 less L2->SM operand traffic than a single-CTA 128 x BN tile — the limit this
// engine was running into.
//   warp 0      TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1      MMA issuer    (one elected thread of the leader CTA, tcgen05.mma, 128|256 x BN x 16|8 per instruction)
//   warps 2..9  epilogue      (tcgen05.ld 32x32b, fused bias / GELU / dGELU / column sums; per-WARP staging + TMA stores)
// Two TMEM accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.
// Operands may be K-major (x @ W^T forward) or MN-major (wgrad reads dy and x as [K=tokens, M|N]); only the TMA box
// pattern and the smem-descriptor strides differ.
// Split-K (wgrad: K = all tokens) accumulates with vector fp32 reductions (red.global.add.v4.f32).
#include "common.cuh"
#include "sm100.cuh"

#include <cstdlib>
#include <cstring>
#include <mutex>

namespace simseg {

using namespace sm100;

constexpr int kBM = 128;            // tile rows per CTA = TMEM lanes
constexpr int kSwizzleBytes = 128;  // one swizzle atom row = one K block (K-major) / 64 MN elements (bf16)
constexpr int kEpiBoxBytes = 32 * 64;   // epilogue staging box: 32 rows x 32 bf16 columns, SWIZZLE_64B (TMA store/load box)
constexpr int kMaxSmem = 232448;

// internal epilogues of the similarity family (fp32 scores never leave the SM; see sim_loss.cu for the entry points)
constexpr int kEpiNceFwd = 16;    // per-(row, tile-half) online-softmax partials of acc / clamp(temp) + target logit
constexpr int kEpiNceBwd = 17;    // G = gs * (softmax - onehot) / t as bf16 hi | lo operands; dtemp
constexpr int kEpiRank = 18;      // rank[row] += #{col : acc beats the row's best matching score}
constexpr int kEpiBest = 19;      // bestkey[row] = max over same-group columns of (ordered score, lowest column)
__host__ __device__ constexpr bool is_sim_epi(int e) { return e >= kEpiNceFwd; }

struct GemmParams {
  // ---- similarity family ----------------------------------------------------------------------------------------
  int32_t split3, kb_seg;          // split-bf16 operands: k-blocks [0,kb_seg) = Ahi*Bhi, then Ahi*Blo, then Alo*Bhi
  const float* temperature;        // device scalar (clamped to [0.001, 0.5] at use)
  int32_t row_offset;              // target column of row i = row_offset + i
  float4* part;                    // NCE fwd: [M][part_ld] (max, sumexp, best, bitcast argmax)
  int32_t part_ld;
  float* zt;                       // NCE fwd: [M] target logit
  const float* lse;                // NCE bwd: [M]
  float grad_scale;
  float* dtemp;
  int64_t lo_off;                  // NCE bwd: G lo half at d + lo_off (bf16 elements)
  const float* best;               // rank: [M] best matching score, bestj [M] its column (-1: none)
  const int32_t* bestj;
  const int64_t* lgid;
  const int64_t* rgid;
  int32_t* rank;
  unsigned long long* bestkey;     // kEpiBest: [M] atomicMax keys
  const int32_t* tile_list;        // optional: only these mn tiles are visited (count in *tile_count)
  const int32_t* tile_count;
  // ---------------------------------------------------------------------------------------------------------------
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
  int32_t tma_epi;         // bf16 output through per-warp swizzled smem staging + TMA store (coalesced), else direct stores
  void* aux2;              // DGELU only: gelu(aux) written next to the gradient (bf16 [M,N])
  int32_t a_3d, b_3d;      // MN-major operand loaded with one 3-D TMA box per k-block
  int32_t dbg;             // bench-only: 1 = no TMA after the ring is primed, 2 = no MMA (results are garbage)
  int32_t d_trans;         // fp32 atomic output stored transposed: element (m, n) at d[n * ldd + m]  (wgrad computed as dW^T)
};

// NEW = number of epilogue warps: 8 (two per TMEM lane quarter), or 16 for the activation epilogues (GELU / dGELU), whose
// per-element work (MUFU + a dozen FMA-pipe instructions in dependent chains) needs four warps per scheduler to fill the
// issue slots — with two, the epilogue of a 128 x 256 tile took ~2x the tile's MMA time (profiles/r01_ncu_gemm.txt).
template <int BN, int EPI, int CTAS, int NEW = 8>
struct GemmCfg {
  static constexpr int kEpiWarps = NEW;
  static constexpr int kThreads = 32 * (2 + NEW);
  static constexpr int kGroups = NEW / 4;                      // warps per TMEM lane quarter = column groups of a tile
  static_assert(NEW == 8 || NEW == 16, "8 or 16 epilogue warps");
  static constexpr bool kGroupsOk = (BN / 32) % kGroups == 0 || BN > 256;   // 32-column chunks divide evenly among the groups
  static constexpr int kABytes = kBM * kSwizzleBytes;          // 16 KB: this CTA's 128 rows of A, one k-block
  static constexpr int kBRows = BN / CTAS;                     // B columns staged by this CTA
  static constexpr int kBBytes = kBRows * kSwizzleBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  // per-warp epilogue staging, 8 warps: EPI_NONE rotates 2 boxes; BIAS_GELU 2 x (out, aux); DGELU 3 aux-in boxes + 1 out box
  //                             16 warps: BIAS_GELU (out, aux) single-buffered; DGELU 2 aux-in boxes + 1 out box
  static constexpr bool kAct = (EPI == SIMSEG_EPI_BIAS_GELU || EPI == SIMSEG_EPI_DGELU);
  static constexpr int kStgBoxes = (BN > 256) ? 0 : (NEW == 16 ? (EPI == SIMSEG_EPI_DGELU ? 3 : 2) : (kAct ? 4 : 2));
  static constexpr int kStgPerWarp = kStgBoxes * kEpiBoxBytes;
  static constexpr int kStagingBytes = NEW * kStgPerWarp;
  static constexpr int kBarBytes = 1024;
  static constexpr int kStagesFit = (kMaxSmem - 1024 - kStagingBytes - kBarBytes) / kStageBytes;
  static constexpr int kStages = kStagesFit > 8 ? 8 : kStagesFit;
  // BN <= 256: two accumulator stages (epilogue of tile i overlaps the MMAs of tile i+1).  BN = 384 / 512 ("wide", CTA
  // pairs, split-K wgrad): ONE accumulator over up to all 512 TMEM columns, two MMAs (N = 256 + BN-256) per k-step —
  // each unit works on one long K slice, so there is nothing to overlap, and the operand bytes per flop drop to
  // (32 KB + BN*128 B) per 256 x BN x 64 MACs.
  static constexpr bool kWide = BN > 256;
  static constexpr int kAccStages = kWide ? 1 : 2;
  static constexpr int kAccStride = (BN <= 128) ? 128 : (kWide ? 512 : 256);    // TMEM columns between accumulators
  static constexpr int kTmemCols = kWide ? 512 : 2 * kAccStride;                // 256 or 512 (power of two)
  static_assert(!kWide || CTAS == 2, "wide tiles are a CTA-pair configuration");
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 /*align slack*/ + kBarBytes;
  static constexpr bool kFits = kStages >= 3 && kGroupsOk;       // checked where the kernel is instantiated
};

// ------------------------------------------------------------------------------------------------
// Coalesced epilogue for bf16 outputs.  Each epilogue warp owns one TMEM lane quarter (32 rows) and every second
// 32-column chunk of the tile; per chunk it goes TMEM -> registers -> fused math -> its OWN SWIZZLE_64B staging box
// [32 rows][32 bf16] -> TMA store.  Nothing is shared between warps, so the only synchronisation is __syncwarp and the
// warp's own bulk-async groups / mbarriers (no CTA-wide barrier per chunk).  The next chunk's tcgen05.ld is in flight
// while the current one is processed, and the accumulator stage is released as soon as the last load has landed.
//   EPI_NONE      out = acc (+bias)                                   2 staging boxes
//   EPI_BIAS_GELU aux = bf16(acc+bias) ; out = gelu(aux)              2 x (aux, out) boxes
//   EPI_DGELU     out = acc * gelu'(aux) ; aux2 = gelu(aux) ; column sums of out
//                 aux boxes are TMA-LOADED two chunks ahead into a 3-slot ring; gelu(aux) overwrites the box in place
//                 and is TMA-stored from there (the backward pass never runs a separate GELU recompute kernel);
//                 column sums stay in registers while the CTA keeps working on the same n-tile.
template <int BN, int EPI, int CTAS>
__device__ __forceinline__ void epilogue_tma(const GemmParams& p, const CUtensorMap& tmap_d, const CUtensorMap& tmap_x,
                                             const CUtensorMap& tmap_x2, uint8_t* stg, uint64_t* acc_full,
                                             uint32_t acc_empty_addr, uint64_t* aux_full, uint32_t tmem_base, int first_tile,
                                             int tile_stride, int total_tiles, int tiles_mn, int row_base, int warp, int lane) {
  using Cfg = GemmCfg<BN, EPI, CTAS>;
  constexpr int kCw = BN / 64;                                      // 32-column chunks per warp and tile
  const int quarter = warp & 3;                                     // TMEM lane quarter this warp may access
  const int half = (warp - 2) >> 2;                                 // chunk parity handled by this warp
  const int sw = (lane >> 1) & 3;                                   // SWIZZLE_64B: 16-byte chunk ^= (row >> 1) & 3
  const uint32_t row_off = static_cast<uint32_t>(lane) * 64;
  const bool has_bias = p.bias != nullptr;
  const bool want_cs = (EPI == SIMSEG_EPI_DGELU) && p.col_sum != nullptr;
  uint32_t f = 0;                                                   // flat chunk counter over (tile, chunk)
  float cs[kCw];
  int cs_n0 = -1;
#pragma unroll
  for (int j = 0; j < kCw; ++j) cs[j] = 0.f;

  auto coords = [&](uint32_t ff, int& row0, int& col0) -> bool {
    const int tile = first_tile + static_cast<int>(ff / kCw) * tile_stride;
    if (tile >= total_tiles) return false;
    const int mn = tile % tiles_mn;
    row0 = (mn / p.n_tiles) * (kBM * CTAS) + row_base + quarter * 32;
    col0 = (mn % p.n_tiles) * BN + (half + 2 * static_cast<int>(ff % kCw)) * 32;
    return true;
  };
  auto flush_cs = [&]() {
    if (cs_n0 < 0) return;
#pragma unroll
    for (int j = 0; j < kCw; ++j) {
      const int c = cs_n0 + (half + 2 * j) * 32 + lane;
      if (c < p.N) atomicAdd(p.col_sum + c, cs[j]);
      cs[j] = 0.f;
    }
  };
  if (EPI == SIMSEG_EPI_DGELU && lane == 0) {
    for (uint32_t ff = 0; ff < 2; ++ff) {
      int r0, c0;
      if (coords(ff, r0, c0)) {
        mbar_arrive_expect_tx(&aux_full[ff % 3], kEpiBoxBytes);
        tma_load_2d(stg + (ff % 3) * kEpiBoxBytes, &tmap_x, &aux_full[ff % 3], c0, r0);
      }
    }
  }

  int acc = 0;
  uint32_t acc_phase = 0;
  for (int tile = first_tile; tile < total_tiles; tile += tile_stride) {
    const int mn = tile % tiles_mn;
    const int n0 = (mn % p.n_tiles) * BN;
    const int row0 = (mn / p.n_tiles) * (kBM * CTAS) + row_base + quarter * 32;
    float bias_r[kCw];                                              // lane l holds bias[column l of chunk j]
#pragma unroll
    for (int j = 0; j < kCw; ++j) {
      const int c = n0 + (half + 2 * j) * 32 + lane;
      bias_r[j] = (has_bias && c < p.N) ? __ldg(p.bias + c) : 0.f;
    }
    if (want_cs && n0 != cs_n0) {
      flush_cs();
      cs_n0 = n0;
    }
    mbar_wait(&acc_full[acc], acc_phase);
    tc_fence_after();
    const uint32_t t_row = tmem_base + acc * Cfg::kAccStride + (static_cast<uint32_t>(quarter * 32) << 16);
    uint32_t ra[32], rb[32];
    tmem_ld_32x32(t_row + half * 32, ra);
#pragma unroll
    for (int j = 0; j < kCw; ++j, ++f) {
      uint32_t(&cur)[32] = (j & 1) ? rb : ra;
      uint32_t(&nxt)[32] = (j & 1) ? ra : rb;
      tmem_ld_wait();
      if (j + 1 < kCw) {
        tmem_ld_32x32(t_row + (half + 2 * (j + 1)) * 32, nxt);
      } else {
        // every TMEM read of this warp has landed in registers: hand the accumulator stage back to the MMA issuer
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(acc_empty_addr + acc * 8);
      }
      const int col = n0 + (half + 2 * j) * 32;
      float v[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(cur[k]);
      if (has_bias) {
        if (col + 32 <= p.N) {
          // every lane reads the same 128 bytes: eight broadcast loads served by L1 (cheaper than 32 shuffles)
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float4 bb = __ldg(b4 + k);
            f2_unpack(f2_add(f2_pack(v[4 * k], v[4 * k + 1]), f2_pack(bb.x, bb.y)), v[4 * k], v[4 * k + 1]);
            f2_unpack(f2_add(f2_pack(v[4 * k + 2], v[4 * k + 3]), f2_pack(bb.z, bb.w)), v[4 * k + 2], v[4 * k + 3]);
          }
        } else {
#pragma unroll
          for (int k = 0; k < 32; ++k) v[k] += __shfl_sync(0xffffffffu, bias_r[j], k);
        }
      }
      uint8_t* so;                                                  // staging box of `out`
      uint8_t* sx;                                                  // staging box of aux (GELU: out; DGELU: in, then aux2 out)
      if (EPI == SIMSEG_EPI_BIAS_GELU) { sx = stg + (f & 1) * (2 * kEpiBoxBytes); so = sx + kEpiBoxBytes; }
      else if (EPI == SIMSEG_EPI_DGELU) { sx = stg + (f % 3) * kEpiBoxBytes; so = stg + 3 * kEpiBoxBytes; }
      else { so = stg + (f & 1) * kEpiBoxBytes; sx = so; }
      so += row_off;
      sx += row_off;
      uint32_t xo[16];                                              // second bf16 output of this row chunk (aux / aux2)
      if (EPI == SIMSEG_EPI_BIAS_GELU) {
#pragma unroll
        for (int k = 0; k < 32; k += 2) {
          xo[k >> 1] = pack_bf16(v[k], v[k + 1]);                   // pre-activation, bf16 — exactly what backward reads
          f2_unpack(gelu_erf_x2(xo[k >> 1]), v[k], v[k + 1]);       // two columns per instruction (FFMA2), one MUFU each
        }
      } else if (EPI == SIMSEG_EPI_DGELU) {
        mbar_wait(&aux_full[f % 3], (f / 3) & 1);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const uint4 u = *reinterpret_cast<const uint4*>(sx + ((q4 ^ sw) << 4));
          const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            f32x2 act, grad;
            gelu_erf_both_x2(w[e], act, grad);
            float a0, a1;
            f2_unpack(act, a0, a1);
            f2_unpack(f2_mul(f2_pack(v[8 * q4 + 2 * e], v[8 * q4 + 2 * e + 1]), grad), v[8 * q4 + 2 * e], v[8 * q4 + 2 * e + 1]);
            xo[4 * q4 + e] = pack_bf16(a0, a1);
          }
        }
      }
      uint32_t o[16];
#pragma unroll
      for (int k = 0; k < 32; k += 2) o[k >> 1] = pack_bf16(v[k], v[k + 1]);
      if (want_cs) {
        // column sums over this warp's 32 rows (butterfly transpose-reduce): lane l ends up with column l
#pragma unroll
        for (int o2 = 16; o2 >= 1; o2 >>= 1) {
          const bool upper = (lane & o2) != 0;
#pragma unroll
          for (int k = 0; k < o2; ++k) {
            const float mine = upper ? v[k + o2] : v[k];
            const float send = upper ? v[k] : v[k + o2];
            v[k] = mine + __shfl_xor_sync(0xffffffffu, send, o2);
          }
        }
        cs[j] += v[0];
      }
      // ---- the staging boxes must have been read by the TMA stores that last used them
      if (lane == 0) {
        if (EPI == SIMSEG_EPI_DGELU) {
          tma_store_wait_read<0>();
          int r2, c2;
          if (coords(f + 2, r2, c2)) {                               // slot (f+2)%3 == (f-1)%3: just released
            mbar_arrive_expect_tx(&aux_full[(f + 2) % 3], kEpiBoxBytes);
            tma_load_2d(stg + ((f + 2) % 3) * kEpiBoxBytes, &tmap_x, &aux_full[(f + 2) % 3], c2, r2);
          }
        } else {
          tma_store_wait_read<1>();
        }
      }
      __syncwarp();
      const bool two = (EPI == SIMSEG_EPI_BIAS_GELU && p.aux != nullptr) || (EPI == SIMSEG_EPI_DGELU && p.aux2 != nullptr);
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const uint32_t off = (q4 ^ sw) << 4;
        *reinterpret_cast<uint4*>(so + off) = make_uint4(o[4 * q4], o[4 * q4 + 1], o[4 * q4 + 2], o[4 * q4 + 3]);
        if (two) *reinterpret_cast<uint4*>(sx + off) = make_uint4(xo[4 * q4], xo[4 * q4 + 1], xo[4 * q4 + 2], xo[4 * q4 + 3]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&tmap_d, so - row_off, col, row0);
        if (EPI == SIMSEG_EPI_BIAS_GELU && p.aux != nullptr) tma_store_2d(&tmap_x, sx - row_off, col, row0);
        if (EPI == SIMSEG_EPI_DGELU && p.aux2 != nullptr) tma_store_2d(&tmap_x2, sx - row_off, col, row0);
        tma_store_commit();
      }
    }
    if (++acc == Cfg::kAccStages) { acc = 0; acc_phase ^= 1; }
  }
  if (want_cs) flush_cs();
  if (lane == 0) tma_store_wait<0>();
}

// ------------------------------------------------------------------------------------------------
// Activation epilogues (BIAS_GELU, DGELU) on SIXTEEN warps: four per TMEM lane quarter, warp (quarter, grp) takes the
// 32-column chunks grp, grp + 4, ... of the tile.  The math per element is the same as in epilogue_tma; what changes is how
// its latency is hidden: not by software-pipelining inside a warp (next chunk's tcgen05.ld in flight, double-buffered
// staging — 140-170 registers) but by four warps per scheduler, each a plain load -> math -> stage -> TMA-store sequence that
// fits 112 registers.  Staging per warp: GELU one (aux, out) pair; DGELU two aux-in boxes (next chunk's pre-activation is
// TMA-loaded while the current one is processed) + one out box.  A box is rewritten only after the bulk-group that last read
// it has finished reading (cp.async.bulk.wait_group.read 0, issued a whole chunk later: it does not stall).
template <int BN, int EPI, int CTAS>
__device__ __forceinline__ void epilogue_act16(const GemmParams& p, const CUtensorMap& tmap_d, const CUtensorMap& tmap_x,
                                               const CUtensorMap& tmap_x2, uint8_t* stg, uint64_t* acc_full,
                                               uint32_t acc_empty_addr, uint64_t* aux_full, uint32_t tmem_base, int first_tile,
                                               int tile_stride, int total_tiles, int tiles_mn, int row_base, int warp, int lane) {
  using Cfg = GemmCfg<BN, EPI, CTAS, 16>;
  constexpr int kG = 4;
  constexpr int kCw = BN / 32 / kG;                                 // chunks per warp and tile (2 for BN = 256)
  const int quarter = warp & 3;
  const int grp = (warp - 2) >> 2;
  const int sw = (lane >> 1) & 3;                                   // SWIZZLE_64B: 16-byte chunk ^= (row >> 1) & 3
  const uint32_t row_off = static_cast<uint32_t>(lane) * 64;
  const bool has_bias = p.bias != nullptr;
  const bool want_cs = (EPI == SIMSEG_EPI_DGELU) && p.col_sum != nullptr;
  uint32_t f = 0;                                                   // flat chunk counter over (tile, chunk)
  float cs[kCw];
  int cs_n0 = -1;
#pragma unroll
  for (int j = 0; j < kCw; ++j) cs[j] = 0.f;

  auto coords = [&](uint32_t ff, int& row0, int& col0) -> bool {
    const int tile = first_tile + static_cast<int>(ff / kCw) * tile_stride;
    if (tile >= total_tiles) return false;
    const int mn = tile % tiles_mn;
    row0 = (mn / p.n_tiles) * (kBM * CTAS) + row_base + quarter * 32;
    col0 = (mn % p.n_tiles) * BN + (grp + kG * static_cast<int>(ff % kCw)) * 32;
    return true;
  };
  auto flush_cs = [&]() {
    if (cs_n0 < 0) return;
#pragma unroll
    for (int j = 0; j < kCw; ++j) {
      const int c = cs_n0 + (grp + kG * j) * 32 + lane;
      if (c < p.N) atomicAdd(p.col_sum + c, cs[j]);
      cs[j] = 0.f;
    }
  };
  if (EPI == SIMSEG_EPI_DGELU && lane == 0) {
    int r0, c0;
    if (coords(0, r0, c0)) {
      mbar_arrive_expect_tx(&aux_full[0], kEpiBoxBytes);
      tma_load_2d(stg, &tmap_x, &aux_full[0], c0, r0);
    }
  }

  int acc = 0;
  uint32_t acc_phase = 0;
  for (int tile = first_tile; tile < total_tiles; tile += tile_stride) {
    const int mn = tile % tiles_mn;
    const int n0 = (mn % p.n_tiles) * BN;
    const int row0 = (mn / p.n_tiles) * (kBM * CTAS) + row_base + quarter * 32;
    if (want_cs && n0 != cs_n0) {
      flush_cs();
      cs_n0 = n0;
    }
    mbar_wait(&acc_full[acc], acc_phase);
    tc_fence_after();
    const uint32_t t_row = tmem_base + acc * Cfg::kAccStride + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll
    for (int j = 0; j < kCw; ++j, ++f) {
      const int col = n0 + (grp + kG * j) * 32;
      // dGELU with gelu(h) kept from the forward (aux2 == NULL): the pre-activation box of the NEXT chunk is requested now, a
      // whole chunk ahead — its slot was last read (generic proxy, fenced) in chunk f - 1.  With the re-emit the slot is the
      // source of chunk f - 1's gelu(h) store and can only be refilled once that store has read it (below).
      const bool early_aux = EPI == SIMSEG_EPI_DGELU && p.aux2 == nullptr && !(p.dbg & 4);
      if (early_aux && lane == 0) {
        int r2, c2;
        if (coords(f + 1, r2, c2)) {
          mbar_arrive_expect_tx(&aux_full[(f + 1) & 1], kEpiBoxBytes);
          tma_load_2d(stg + ((f + 1) & 1) * kEpiBoxBytes, &tmap_x, &aux_full[(f + 1) & 1], c2, r2);
        }
      }
      uint32_t r[32];
      tmem_ld_32x32(t_row + (grp + kG * j) * 32, r);
      tmem_ld_wait();
      if (j + 1 == kCw) {
        // every TMEM read of this warp has landed in registers: hand the accumulator stage back to the MMA issuer
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(acc_empty_addr + acc * 8);
      }
      float v[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(r[k]);
      if (has_bias) {
        if (col + 32 <= p.N) {
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float4 bb = __ldg(b4 + k);
            f2_unpack(f2_add(f2_pack(v[4 * k], v[4 * k + 1]), f2_pack(bb.x, bb.y)), v[4 * k], v[4 * k + 1]);
            f2_unpack(f2_add(f2_pack(v[4 * k + 2], v[4 * k + 3]), f2_pack(bb.z, bb.w)), v[4 * k + 2], v[4 * k + 3]);
          }
        } else {
#pragma unroll
          for (int k = 0; k < 32; ++k) v[k] += (col + k < p.N) ? __ldg(p.bias + col + k) : 0.f;
        }
      }
      uint8_t* sx;                                                  // aux box (GELU: aux out; DGELU: aux in, then aux2 out)
      uint8_t* so;                                                  // out box
      if (EPI == SIMSEG_EPI_BIAS_GELU) { sx = stg; so = stg + kEpiBoxBytes; }
      else { sx = stg + (f & 1) * kEpiBoxBytes; so = stg + 2 * kEpiBoxBytes; }
      sx += row_off;
      so += row_off;
      uint32_t xo[16];
      if (EPI == SIMSEG_EPI_BIAS_GELU) {
#pragma unroll
        for (int k = 0; k < 32; k += 2) {
          xo[k >> 1] = pack_bf16(v[k], v[k + 1]);                   // pre-activation, bf16 — exactly what backward reads
          f2_unpack(gelu_erf_x2(xo[k >> 1]), v[k], v[k + 1]);
        }
      } else {
        mbar_wait(&aux_full[f & 1], (f >> 1) & 1);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const uint4 u = *reinterpret_cast<const uint4*>(sx + ((q4 ^ sw) << 4));
          const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            f32x2 act, grad;
            gelu_erf_both_x2(w[e], act, grad);
            float a0, a1;
            f2_unpack(act, a0, a1);
            f2_unpack(f2_mul(f2_pack(v[8 * q4 + 2 * e], v[8 * q4 + 2 * e + 1]), grad), v[8 * q4 + 2 * e], v[8 * q4 + 2 * e + 1]);
            xo[4 * q4 + e] = pack_bf16(a0, a1);
          }
        }
      }
      uint32_t o[16];
#pragma unroll
      for (int k = 0; k < 32; k += 2) o[k >> 1] = pack_bf16(v[k], v[k + 1]);
      if (want_cs) {
        // column sums over this warp's 32 rows (butterfly transpose-reduce): lane l ends up with column l
#pragma unroll
        for (int o2 = 16; o2 >= 1; o2 >>= 1) {
          const bool upper = (lane & o2) != 0;
#pragma unroll
          for (int k = 0; k < o2; ++k) {
            const float mine = upper ? v[k + o2] : v[k];
            const float send = upper ? v[k] : v[k + o2];
            v[k] = mine + __shfl_xor_sync(0xffffffffu, send, o2);
          }
        }
        cs[j] += v[0];
      }
      // ---- the boxes written below must have been read by the bulk group of the previous chunk
      if (lane == 0) {
        tma_store_wait_read<0>();
        if (EPI == SIMSEG_EPI_DGELU && !early_aux) {
          int r2, c2;
          if (coords(f + 1, r2, c2)) {                               // slot (f+1)&1 == (f-1)&1: its aux2 store has been read
            mbar_arrive_expect_tx(&aux_full[(f + 1) & 1], kEpiBoxBytes);
            tma_load_2d(stg + ((f + 1) & 1) * kEpiBoxBytes, &tmap_x, &aux_full[(f + 1) & 1], c2, r2);
          }
        }
      }
      __syncwarp();
      const bool two = (EPI == SIMSEG_EPI_BIAS_GELU && p.aux != nullptr) || (EPI == SIMSEG_EPI_DGELU && p.aux2 != nullptr);
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const uint32_t off = (q4 ^ sw) << 4;
        *reinterpret_cast<uint4*>(so + off) = make_uint4(o[4 * q4], o[4 * q4 + 1], o[4 * q4 + 2], o[4 * q4 + 3]);
        if (two) *reinterpret_cast<uint4*>(sx + off) = make_uint4(xo[4 * q4], xo[4 * q4 + 1], xo[4 * q4 + 2], xo[4 * q4 + 3]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&tmap_d, so - row_off, col, row0);
        if (EPI == SIMSEG_EPI_BIAS_GELU && p.aux != nullptr) tma_store_2d(&tmap_x, sx - row_off, col, row0);
        if (EPI == SIMSEG_EPI_DGELU && p.aux2 != nullptr) tma_store_2d(&tmap_x2, sx - row_off, col, row0);
        tma_store_commit();
      }
    }
    if (++acc == Cfg::kAccStages) { acc = 0; acc_phase ^= 1; }
  }
  if (want_cs) flush_cs();
  if (lane == 0) tma_store_wait<0>();
}

// ------------------------------------------------------------------------------------------------
// Epilogues of the similarity family: the fp32 score tile is consumed straight out of TMEM and never written.
// A thread owns one row (TMEM lane) and half of the tile's columns, so every per-row reduction is thread-local.
//   kEpiNceFwd  NCE.forward (mml_loss.py:56,73-77 + utils/misc.py:462-477): z = acc / clamp(temp); one online-softmax partial
//               (max, sum exp, best value, first-max column) per (row, n-tile, half) + the target logit z[row, row_offset+row]
//   kEpiNceBwd  G = grad_scale * (exp(z - lse) - [col == target]) / t written as bf16 hi | lo (the split operands of the
//               two gradient GEMMs); dtemp += -(1/t) sum G * acc  (inside the clamp range)
//   kEpiRank    EmbANN/RetrievalMetric (tasks/clip/hooks/utils.py:35-42,63-65) without the sort: rank[row] += number of
//               columns of another group whose score beats the row's best matching score (ties: lower column first)
template <int BN, int EPI>
__device__ __forceinline__ void sim_epilogue_tile(const GemmParams& p, uint32_t t_row, int half, int64_t row, bool row_ok,
                                                  int n0, int nt, int lane) {
  constexpr int kColsPerWarp = BN / 2;
  constexpr int kChunks = kColsPerWarp / 32;
  constexpr float kLog2e = 1.44269504088896341f;
  float traw = 0.02f;
  if (EPI != kEpiRank) traw = __ldg(p.temperature);
  const float t = fminf(fmaxf(traw, 0.001f), 0.5f);
  const float inv_t = 1.0f / t;
  const int64_t tgt = p.row_offset + row;
  // ---- per-row state
  float m2 = -INFINITY, l = 0.f, best = -INFINITY, ztv = 0.f;     // NCE fwd (m2 in the log2 domain)
  int besti = 0x7fffffff;
  bool have_t = false;
  float L2 = 0.f, dsum = 0.f;                                      // NCE bwd
  if (EPI == kEpiNceBwd) L2 = (row_ok ? __ldg(p.lse + row) : 0.f) * kLog2e;
  float thr = 0.f;                                                 // rank
  int js = -1, cnt = 0;
  long long gid = 0;
  if (EPI == kEpiRank && row_ok) { thr = __ldg(p.best + row); js = __ldg(p.bestj + row); gid = __ldg(p.lgid + row); }
  unsigned long long key = 0ull;                                   // best
  if (EPI == kEpiBest && row_ok) gid = __ldg(p.lgid + row);
#pragma unroll 1
  for (int c = 0; c < kChunks; ++c) {
    const int col_in_tile = half * kColsPerWarp + c * 32;
    const int64_t col0 = n0 + col_in_tile;
    uint32_t r[32];
    tmem_ld_32x32(t_row + col_in_tile, r);
    tmem_ld_wait();
    if (col0 >= p.N) continue;                                     // whole chunk is N padding (warp-uniform)
    const int nvalid = (col0 + 32 <= p.N) ? 32 : static_cast<int>(p.N - col0);
    if (EPI == kEpiNceFwd) {
      const float sc = inv_t * kLog2e;
      float cm = -INFINITY;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float s = __uint_as_float(r[j]);
        if (j < nvalid) {
          if (s > best) { best = s; besti = static_cast<int>(col0) + j; }      // first maximum (ascending columns)
          cm = fmaxf(cm, s);
          if (col0 + j == tgt) { ztv = s * inv_t; have_t = true; }
        }
      }
      cm *= sc;
      if (cm > m2) { l *= exp2f(m2 - cm); m2 = cm; }
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float e;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(__uint_as_float(r[j]), sc, -m2)));
        acc += (j < nvalid) ? e : 0.f;
      }
      l += acc;
    } else if (EPI == kEpiNceBwd) {
      const float sc = inv_t * kLog2e;
      const float gs = p.grad_scale * inv_t;
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        float g[2];
#pragma unroll
        for (int e2 = 0; e2 < 2; ++e2) {
          const float s = __uint_as_float(r[j + e2]);
          float pr;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pr) : "f"(fmaf(s, sc, -L2)));
          if (col0 + j + e2 == tgt) pr -= 1.0f;
          g[e2] = (row_ok && j + e2 < nvalid) ? gs * pr : 0.f;
          dsum = fmaf(g[e2], s, dsum);
        }
        const uint32_t h = pack_bf16(g[0], g[1]);
        hi[j >> 1] = h;
        lo[j >> 1] = pack_bf16(g[0] - bf16_lo(h), g[1] - bf16_hi(h));
      }
      if (row_ok) {
        __nv_bfloat16* dh = reinterpret_cast<__nv_bfloat16*>(p.d) + row * p.ldd + col0;
        __nv_bfloat16* dl = dh + p.lo_off;
        if (nvalid == 32) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            reinterpret_cast<uint4*>(dh)[q] = make_uint4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
            reinterpret_cast<uint4*>(dl)[q] = make_uint4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (j < nvalid) {
              const uint32_t h = hi[j >> 1], w = lo[j >> 1];
              reinterpret_cast<uint16_t*>(dh)[j] = static_cast<uint16_t>((j & 1) ? (h >> 16) : (h & 0xffffu));
              reinterpret_cast<uint16_t*>(dl)[j] = static_cast<uint16_t>((j & 1) ? (w >> 16) : (w & 0xffffu));
            }
          }
        }
      }
    } else if (EPI == kEpiBest) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (j < nvalid) {
          const int col = static_cast<int>(col0) + j;
          if (row_ok && __ldg(reinterpret_cast<const long long*>(p.rgid) + col) == gid) {
            uint32_t u = r[j];
            u ^= (u >> 31) ? 0xffffffffu : 0x80000000u;                        // monotonic float -> uint
            const unsigned long long k2 = (static_cast<unsigned long long>(u) << 32) | (0xffffffffu - static_cast<uint32_t>(col));
            key = k2 > key ? k2 : key;
          }
        }
      }
    } else {   // kEpiRank
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (j < nvalid) {
          const float s = __uint_as_float(r[j]);
          const int col = static_cast<int>(col0) + j;
          const bool beats = (s > thr) || (s == thr && col < js);
          const bool same = __ldg(reinterpret_cast<const long long*>(p.rgid) + col) == gid;    // warp-uniform address
          cnt += (beats && !same) ? 1 : 0;
        }
      }
    }
  }
  if (EPI == kEpiNceFwd) {
    if (row_ok) {
      constexpr float kLn2 = 0.69314718055994531f;
      p.part[row * p.part_ld + nt * 2 + half] = make_float4(m2 * kLn2, l, best * inv_t, __int_as_float(besti));
      if (have_t) p.zt[row] = ztv;
    }
  } else if (EPI == kEpiNceBwd) {
    if (p.dtemp != nullptr) {
      dsum = warp_sum(dsum);
      if (lane == 0 && dsum != 0.f && traw >= 0.001f && traw <= 0.5f) atomicAdd(p.dtemp, -dsum * inv_t);
    }
  } else if (EPI == kEpiBest) {
    if (key) atomicMax(p.bestkey + row, key);
  } else {
    if (row_ok && js >= 0 && cnt) atomicAdd(p.rank + row, cnt);
  }
}

// ------------------------------------------------------------------------------------------------
template <int BN, int EPI, int CTAS, int NEW = 8>
__global__ void __launch_bounds__(32 * (2 + NEW), 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
            const __grid_constant__ CUtensorMap tmap_d, const __grid_constant__ CUtensorMap tmap_x,
            const __grid_constant__ CUtensorMap tmap_x2, const GemmParams p) {
  using Cfg = GemmCfg<BN, EPI, CTAS, NEW>;
  static_assert(Cfg::kFits, "operand ring too shallow, or chunks do not divide among the epilogue warps");
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment is required by SWIZZLE_128B (descriptor base_offset = 0); the dynamic-smem base offset is the
  // same in both CTAs of a pair, so every carved address below is at the same offset in the peer.
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
  uint8_t* stg = smem + Cfg::kStages * Cfg::kStageBytes;      // per-warp epilogue staging boxes
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg + Cfg::kStagingBytes);
  uint64_t* full_bar = bars;                       // [8]   TMA -> MMA            (leader's are used by a pair)
  uint64_t* empty_bar = bars + 8;                  // [8]   MMA -> TMA            (multicast to both CTAs of a pair)
  uint64_t* acc_full = bars + 16;                  // [2]   MMA -> epilogue       (multicast)
  uint64_t* acc_empty = bars + 18;                 // [2]   epilogue -> MMA       (leader's; both CTAs' warps arrive)
  uint64_t* aux_full = bars + 20;                  // [NEW warps][3]  TMA (aux-in boxes of the DGELU epilogue) -> warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20 + 3 * 16);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = (CTAS == 2) ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], NEW * CTAS);
    }
    for (int s = 0; s < 3 * NEW; ++s) mbar_init(&aux_full[s], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CTAS == 2) { tmem_alloc_2sm(tmem_slot, Cfg::kTmemCols); tmem_relinquish_2sm(); }
    else { tmem_alloc(tmem_slot, Cfg::kTmemCols); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();               // the peer's barriers exist before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = p.tile_list ? __ldg(p.tile_count) : p.m_tiles * p.n_tiles * p.splits;
  const int tiles_mn = p.m_tiles * p.n_tiles;
  const int k_elems = kSwizzleBytes / p.elem_bytes;      // K elements per k-block: 64 (bf16) or 32 (tf32)
  const int first_tile = static_cast<int>(blockIdx.x) / CTAS;
  const int tile_stride = static_cast<int>(gridDim.x) / CTAS;
  const int row_base = static_cast<int>(rank) * kBM;     // this CTA's rows inside the (pair's) tile

  if (warp == 0) {
    // =============================== TMA producer ===============================
    // The whole warp runs the loop (so every value stays in uniform registers, which is what UTMALDG wants); one
    // elected lane issues.
    {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t full0 = (CTAS == 2) ? mapa_shared(smem_u32(&full_bar[0]), 0) : 0u;   // leader's full barriers
      const int mn_atom = kSwizzleBytes / p.elem_bytes;              // 64 bf16 / 32 tf32 elements
      for (int tile = first_tile; tile < total_tiles; tile += tile_stride) {
        const int split = tile / tiles_mn;
        const int mn = p.tile_list ? __ldg(p.tile_list + tile) : tile - split * tiles_mn;
        const int m0 = (mn / p.n_tiles) * (kBM * CTAS) + row_base;
        const int n0 = (mn % p.n_tiles) * BN + static_cast<int>(rank) * Cfg::kBRows;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            uint8_t* sa = smem + stage * Cfg::kStageBytes;
            uint8_t* sb = sa + Cfg::kABytes;
            if (p.dbg == 1 && (phase != 0 || tile != first_tile)) {
              if (leader) mbar_arrive(&full_bar[stage]);
            } else {
              if (leader) mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes * CTAS);
              const uint32_t fb = full0 + stage * 8;
              // split-bf16 products (fp32-grade similarity GEMMs): three K segments over (hi,hi), (hi,lo), (lo,hi)
              const CUtensorMap* pa = &tmap_a;
              const CUtensorMap* pb = &tmap_b;
              int kk = kb;
              if (p.split3) {
                const int seg = kb / p.kb_seg;
                kk = kb - seg * p.kb_seg;
                if (seg == 2) pa = &tmap_x;
                if (seg == 1) pb = &tmap_x2;
              }
              const int k0 = kk * k_elems;
              if (!p.a_mn) {
                if (CTAS == 2) tma_load_2d_2sm(sa, pa, fb, k0, m0);
                else tma_load_2d(sa, pa, &full_bar[stage], k0, m0);
              } else {
                // MN-major: boxes of (swizzle-row of MN elements) x k rows; one 3-D box when the extent allows
                const int nbox = kBM / mn_atom;
                const int box_bytes = Cfg::kABytes / nbox;
                if (p.a_3d) {
                  if (CTAS == 2) tma_load_3d_2sm(sa, pa, fb, 0, k0, m0 / mn_atom);
                  else tma_load_3d(sa, pa, &full_bar[stage], 0, k0, m0 / mn_atom);
                } else {
                  for (int j = 0; j < nbox; ++j) {
                    if (CTAS == 2) tma_load_2d_2sm(sa + j * box_bytes, pa, fb, m0 + j * mn_atom, k0);
                    else tma_load_2d(sa + j * box_bytes, pa, &full_bar[stage], m0 + j * mn_atom, k0);
                  }
                }
              }
              if (!p.b_mn) {
                if (CTAS == 2) tma_load_2d_2sm(sb, pb, fb, k0, n0);
                else tma_load_2d(sb, pb, &full_bar[stage], k0, n0);
              } else if (Cfg::kWide) {
                // wide pair tile: per CTA two column groups (one per MMA), one 3-D box per 64-column atom
                constexpr int kAtoms0 = 2, kAtoms1 = (BN - 256) / 128;      // atoms per CTA for MMA 0 (N=256) and MMA 1
                const int nt0 = (mn % p.n_tiles) * BN;
#pragma unroll
                for (int a2 = 0; a2 < kAtoms0 + kAtoms1; ++a2) {
                  const int col = a2 < kAtoms0 ? nt0 + static_cast<int>(rank) * 128 + a2 * 64
                                               : nt0 + 256 + static_cast<int>(rank) * (kAtoms1 * 64) + (a2 - kAtoms0) * 64;
                  tma_load_3d_2sm(sb + a2 * 8192, pb, fb, 0, k0, col / 64);
                }
              } else {
                const int nbox = Cfg::kBRows / mn_atom;
                const int box_bytes = Cfg::kBBytes / nbox;
                if (p.b_3d) {
                  if (CTAS == 2) tma_load_3d_2sm(sb, pb, fb, 0, k0, n0 / mn_atom);
                  else tma_load_3d(sb, pb, &full_bar[stage], 0, k0, n0 / mn_atom);
                } else {
                  for (int j = 0; j < nbox; ++j) {
                    if (CTAS == 2) tma_load_2d_2sm(sb + j * box_bytes, pb, fb, n0 + j * mn_atom, k0);
                    else tma_load_2d(sb + j * box_bytes, pb, &full_bar[stage], n0 + j * mn_atom, k0);
                  }
                }
              }
            }
          }
          __syncwarp();
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (leader CTA only) ===============================
    // Warp-uniform loop, one elected lane issues: descriptors are one 64-bit add away from per-kernel constants, so a
    // k-block costs a few dozen instructions (the issuing thread, not the tensor pipe, used to be the limit).
    if (leader) {
      const uint32_t idesc = make_idesc(p.elem_bytes == 2 ? 1u : 2u, p.a_mn, p.b_mn, kBM * CTAS, Cfg::kWide ? 256 : BN);
      const uint32_t idesc1 = make_idesc(1u, p.a_mn, p.b_mn, kBM * CTAS, Cfg::kWide ? BN - 256 : 16);   // wide tiles: second MMA
      // K-major: consecutive UMMA_K slices are 32 B apart inside the 128 B swizzle row; 8-row groups 1024 B apart.
      // MN-major: a UMMA_K slice (16 bf16 / 8 tf32 k-rows) spans k-rows of 128 B each,
      //           8-row groups 1024 B apart (SBO), 64-element MN atoms one TMA box apart (LBO).
      const bool is_bf16 = p.elem_bytes == 2;
      const uint32_t umma_k = 32 / p.elem_bytes;                // 16 (bf16) / 8 (tf32)
      const uint32_t a_step = p.a_mn ? umma_k * kSwizzleBytes : 32;
      const uint32_t b_step = p.b_mn ? umma_k * kSwizzleBytes : 32;
      const uint32_t k_rows_bytes = (kSwizzleBytes / p.elem_bytes) * kSwizzleBytes;   // one MN-major box: k_elems rows x 128 B
      const uint64_t adesc0 = make_smem_desc_sw128(smem_u32(smem), p.a_mn ? k_rows_bytes : 16, 1024);
      const uint64_t bdesc0 = make_smem_desc_sw128(smem_u32(smem) + Cfg::kABytes, p.b_mn ? k_rows_bytes : 16, 1024);
      const uint64_t a_inc = a_step >> 4, b_inc = b_step >> 4;   // descriptor start-address field is in 16-byte units
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_stride) {
        const int split = tile / tiles_mn;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::kAccStride;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t ad = adesc0 + static_cast<uint64_t>(stage * (Cfg::kStageBytes >> 4));
            const uint64_t bd = bdesc0 + static_cast<uint64_t>(stage * (Cfg::kStageBytes >> 4));
            const uint32_t first = (kb > kb0) ? 1u : 0u;
            if (!(p.dbg == 2 && kb > kb0)) {
              if (is_bf16) {
                if (Cfg::kWide) {
                  // columns 0..255 and 256..BN-1 of the accumulator: two MMAs per k-step (bf16, MN-major operands)
                  const uint64_t bd1 = bd + (16384 >> 4);                  // second column group: 2 atoms x 8 KB further
#pragma unroll
                  for (int kk = 0; kk < 4; ++kk) {
                    const uint32_t accf = (kk == 0) ? first : 1u;
                    umma_f16_2sm(d_tmem, ad + kk * a_inc, bd + kk * b_inc, idesc, accf);
                    umma_f16_2sm(d_tmem + 256, ad + kk * a_inc, bd1 + kk * b_inc, idesc1, accf);
                  }
                } else if (CTAS == 2) {
                  umma_f16_2sm(d_tmem, ad, bd, idesc, first);
                  umma_f16_2sm(d_tmem, ad + a_inc, bd + b_inc, idesc, 1u);
                  umma_f16_2sm(d_tmem, ad + 2 * a_inc, bd + 2 * b_inc, idesc, 1u);
                  umma_f16_2sm(d_tmem, ad + 3 * a_inc, bd + 3 * b_inc, idesc, 1u);
                } else {
                  umma_f16(d_tmem, ad, bd, idesc, first);
                  umma_f16(d_tmem, ad + a_inc, bd + b_inc, idesc, 1u);
                  umma_f16(d_tmem, ad + 2 * a_inc, bd + 2 * b_inc, idesc, 1u);
                  umma_f16(d_tmem, ad + 3 * a_inc, bd + 3 * b_inc, idesc, 1u);
                }
              } else {
                if (CTAS == 2) {
                  umma_tf32_2sm(d_tmem, ad, bd, idesc, first);
                  umma_tf32_2sm(d_tmem, ad + a_inc, bd + b_inc, idesc, 1u);
                  umma_tf32_2sm(d_tmem, ad + 2 * a_inc, bd + 2 * b_inc, idesc, 1u);
                  umma_tf32_2sm(d_tmem, ad + 3 * a_inc, bd + 3 * b_inc, idesc, 1u);
                } else {
                  umma_tf32(d_tmem, ad, bd, idesc, first);
                  umma_tf32(d_tmem, ad + a_inc, bd + b_inc, idesc, 1u);
                  umma_tf32(d_tmem, ad + 2 * a_inc, bd + 2 * b_inc, idesc, 1u);
                  umma_tf32(d_tmem, ad + 3 * a_inc, bd + 3 * b_inc, idesc, 1u);
                }
              }
            }
            // frees this smem slot (in both CTAs of a pair) once the MMAs have read it
            if (CTAS == 2) umma_commit_2sm(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
            // accumulator complete -> epilogue warps (of both CTAs)
            if (kb == kb1 - 1) { if (CTAS == 2) umma_commit_2sm(&acc_full[acc]); else umma_commit(&acc_full[acc]); }
          }
          __syncwarp();
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        if (++acc == Cfg::kAccStages) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // =============================== epilogue ===============================
    // the leader's acc_empty barriers (a shared::cta address is a valid shared::cluster address of the own CTA)
    const uint32_t acc_empty_addr = (CTAS == 2) ? mapa_shared(smem_u32(&acc_empty[0]), 0) : smem_u32(&acc_empty[0]);
    if (p.tma_epi) {
      if constexpr (NEW == 16)
        epilogue_act16<BN, EPI, CTAS>(p, tmap_d, tmap_x, tmap_x2, stg + (warp - 2) * Cfg::kStgPerWarp, acc_full, acc_empty_addr,
                                      aux_full + 3 * (warp - 2), tmem_base, first_tile, tile_stride, total_tiles, tiles_mn,
                                      row_base, warp, lane);
      else
        epilogue_tma<BN, EPI, CTAS>(p, tmap_d, tmap_x, tmap_x2, stg + (warp - 2) * Cfg::kStgPerWarp, acc_full, acc_empty_addr,
                                    aux_full + 3 * (warp - 2), tmem_base, first_tile, tile_stride, total_tiles, tiles_mn, row_base,
                                    warp, lane);
    } else {
    const int ew = warp - 2;                 // 0..NEW-1
    const int quarter = warp & 3;            // TMEM lane quarter this warp may access
    const int half = ew >> 2;                // column group handled by this warp (two groups = halves with 8 warps)
    constexpr int kColsPerWarp = BN / Cfg::kGroups;
    constexpr int kChunks = kColsPerWarp / 32;
    static_assert(kColsPerWarp % 32 == 0, "BN must be a multiple of 32 x column groups");
    const int row_in_tile = quarter * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = first_tile; tile < total_tiles; tile += tile_stride) {
      const int split = tile / tiles_mn;
      const int mn = p.tile_list ? __ldg(p.tile_list + tile) : tile - split * tiles_mn;
      const int m0 = (mn / p.n_tiles) * (kBM * CTAS) + row_base;
      const int n0 = (mn % p.n_tiles) * BN;
      const int64_t row = m0 + row_in_tile;
      const bool row_ok = row < p.M;
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + acc * Cfg::kAccStride + (static_cast<uint32_t>(quarter * 32) << 16);
      float rs = 1.0f;
      if (EPI == SIMSEG_EPI_ROWSCALE) rs = row_ok ? p.row_scale[row] : 0.0f;
      if (is_sim_epi(EPI)) sim_epilogue_tile<BN, EPI>(p, t_row, half, row, row_ok, n0, mn % p.n_tiles, lane);
#pragma unroll 1
      for (int c = 0; c < (is_sim_epi(EPI) ? 0 : kChunks); ++c) {
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
          if (p.atomic_out && p.d_trans) {
            // D^T: lanes hold consecutive m, so each reduction instruction covers 32 consecutive floats of one output row
            float* dp = reinterpret_cast<float*>(p.d) + col0 * p.ldd + row;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dp + j * p.ldd), "f"(v[j]) : "memory");
          } else if (p.atomic_out) {
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
      if (lane == 0) mbar_arrive_cluster(acc_empty_addr + acc * 8);
      if (++acc == Cfg::kAccStages) { acc = 0; acc_phase ^= 1; }
    }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();               // the peer may still read our smem / signal our barriers until here
  if (warp == 1) {
    tc_fence_after();
    if (CTAS == 2) tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols); else tmem_dealloc(tmem_base, Cfg::kTmemCols);
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
static int make_tmap_sw(CUtensorMap* m, const void* ptr, int elem_bytes, int64_t rows, int64_t cols, int64_t ld,
                        int box_cols, int box_rows, CUtensorMapSwizzle swz);
int make_tmap(CUtensorMap* m, const void* ptr, int elem_bytes, int64_t rows, int64_t cols, int64_t ld,
              int box_cols, int box_rows) {
  return make_tmap_sw(m, ptr, elem_bytes, rows, cols, ld, box_cols, box_rows, CU_TENSOR_MAP_SWIZZLE_128B);
}
static int make_tmap_sw(CUtensorMap* m, const void* ptr, int elem_bytes, int64_t rows, int64_t cols, int64_t ld,
                        int box_cols, int box_rows, CUtensorMapSwizzle swz) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled not available (driver entry point lookup failed)"); return SIMSEG_ERR_CUDA; }
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * elem_bytes};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
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
  static int promo = -1;
  if (promo < 0) { const char* e = getenv("SIMSEG_TMA_PROMO"); promo = e ? atoi(e) : 3; }
  const CUtensorMapL2promotion pr = promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                    : promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  CUresult r = enc(m, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(3d) failed (%d): ptr=%p k_rows=%lld mn=%lld ld=%lld box_k=%d atoms=%d", static_cast<int>(r),
              ptr, static_cast<long long>(k_rows), static_cast<long long>(mn_cols), static_cast<long long>(ld), box_k, box_atoms);
    return SIMSEG_ERR_CUDA;
  }
  return SIMSEG_OK;
}

template <int BN, int EPI, int CTAS, int NEW = 8>
static int launch_gemm(Ctx* ctx, const CUtensorMap* tm, const GemmParams& p, int grid, cudaStream_t st) {
  using Cfg = GemmCfg<BN, EPI, CTAS, NEW>;
  static bool attr_set = false;
  auto kfn = gemm_kernel<BN, EPI, CTAS, NEW>;
  if (!attr_set) {
    SIMSEG_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  if (CTAS == 1) {
    kfn<<<grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(tm[0], tm[1], tm[2], tm[3], tm[4], p);
  } else {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(Cfg::kThreads);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    SIMSEG_CUDA(cudaLaunchKernelEx(&cfg, kfn, tm[0], tm[1], tm[2], tm[3], tm[4], p));
  }
  ctx->launches++;
  SIMSEG_LAUNCH_CHECK();
  return SIMSEG_OK;
}

template <int BN, int CTAS>
static int dispatch_epi(Ctx* ctx, int epi, const CUtensorMap* tm, const GemmParams& p, int grid, cudaStream_t st, bool act16) {
  // activation epilogues on 16 warps (four per scheduler) wherever the larger staging area leaves a >= 3-stage operand
  // ring; `reserved & 64` keeps the 8-warp version for A/B timing
  if constexpr (GemmCfg<BN, SIMSEG_EPI_BIAS_GELU, CTAS, 16>::kFits) {
    if (act16 && epi == SIMSEG_EPI_BIAS_GELU) return launch_gemm<BN, SIMSEG_EPI_BIAS_GELU, CTAS, 16>(ctx, tm, p, grid, st);
  }
  if constexpr (GemmCfg<BN, SIMSEG_EPI_DGELU, CTAS, 16>::kFits) {
    if (act16 && epi == SIMSEG_EPI_DGELU) return launch_gemm<BN, SIMSEG_EPI_DGELU, CTAS, 16>(ctx, tm, p, grid, st);
  }
  switch (epi) {
    case SIMSEG_EPI_NONE: return launch_gemm<BN, SIMSEG_EPI_NONE, CTAS>(ctx, tm, p, grid, st);
    case SIMSEG_EPI_BIAS_GELU: return launch_gemm<BN, SIMSEG_EPI_BIAS_GELU, CTAS>(ctx, tm, p, grid, st);
    case SIMSEG_EPI_BIAS_RESIDUAL: return launch_gemm<BN, SIMSEG_EPI_BIAS_RESIDUAL, CTAS>(ctx, tm, p, grid, st);
    case SIMSEG_EPI_DGELU: return launch_gemm<BN, SIMSEG_EPI_DGELU, CTAS>(ctx, tm, p, grid, st);
    case SIMSEG_EPI_ROWSCALE: return launch_gemm<BN, SIMSEG_EPI_ROWSCALE, CTAS>(ctx, tm, p, grid, st);
  }
  set_error("unknown epilogue %d", epi);
  return SIMSEG_ERR_INVALID;
}

template <int EPI>
static int dispatch_sim(Ctx* ctx, int bn, int ctas, const CUtensorMap* tm, const GemmParams& p, int grid, cudaStream_t st) {
  if (ctas == 2) return launch_gemm<256, EPI, 2>(ctx, tm, p, grid, st);
  if (bn == 128) return launch_gemm<128, EPI, 1>(ctx, tm, p, grid, st);
  return launch_gemm<256, EPI, 1>(ctx, tm, p, grid, st);
}

int gemm_impl(Ctx* ctx, const simseg_gemm_args* a, cudaStream_t st) { return gemm_sim_impl(ctx, a, nullptr, st); }

// `reserved` bits (bench / debug only): 1 = no TMA, 2 = no MMA, 4 = 2-D boxes for MN-major operands, 8 = direct-store
// epilogue, 16 = force single-CTA tiles, 32 = force CTA pairs, 64 = activation epilogues on 8 warps (the round-1 version),
// 128 = no wide (256 x 384|512) split-K tiles, 256 = 16-warp GELU epilogue regardless of K, 512 = dGELU pre-activation
// boxes requested late (after the chunk's math, the order the gelu(h) re-emit needs) even when gelu(h) was kept.
// `sim` != NULL: split-bf16 operands (a / b are the hi halves, sim->a_lo / b_lo the lo halves, K counts ONE segment) and
// optionally one of the similarity epilogues (sim->epi = 0 keeps a->epilogue).
int gemm_sim_impl(Ctx* ctx, const simseg_gemm_args* a, const GemmSim* sim, cudaStream_t st) {
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

  const int mn_atom = kSwizzleBytes / eb;
  const int k_elems = kSwizzleBytes / eb;
  const int kb_seg = static_cast<int>(cdiv(a->K, k_elems));
  const int kb_total = sim ? 3 * kb_seg : kb_seg;
  const int sim_epi = sim ? sim->epi : 0;
  if (sim) {
    SIMSEG_CHECK_ARG(eb == 2 && sim->a_lo && sim->b_lo, "gemm: split operands need bf16 hi/lo pairs");
    SIMSEG_CHECK_ARG((reinterpret_cast<uintptr_t>(sim->a_lo) & 15) == 0 && (reinterpret_cast<uintptr_t>(sim->b_lo) & 15) == 0,
                     "gemm: lo operands must be 16-byte aligned");
    SIMSEG_CHECK_ARG(sim_epi == 0 || (!a->a_major && !a->b_major && a->epilogue == SIMSEG_EPI_NONE),
                     "gemm: similarity epilogues take K-major operands");
  }
  const bool can_split = !sim_epi && a->epilogue == SIMSEG_EPI_NONE && a->out_dtype == SIMSEG_F32 && a->col_sum == nullptr;

  // ---- tile shape: BN in {128,192,256}, single CTA (128 rows) or CTA pair (256 rows)
  int bn = a->tile_n;
  SIMSEG_CHECK_ARG(bn == 0 || bn == 128 || bn == 192 || bn == 256, "gemm: tile_n must be 0/128/192/256 (got %d)", bn);
  int ctas = (a->reserved & 16) ? 1 : ((a->reserved & 32) ? 2 : 0);
  if (eb != 2) ctas = 1;                                       // pairs are only instantiated for the bf16 path
  auto pair_ok = [&](int n) { return !(a->b_major && (n / 2) % mn_atom != 0); };
  if (ctas == 2 && bn != 0) SIMSEG_CHECK_ARG(pair_ok(bn), "gemm: CTA pairs need tile_n/2 to be a whole number of MN atoms");
  if (bn == 0 || ctas == 0) {
    // Measured on B200 (profiles/r01_gemm_microbench_v2.txt):
    //  * K-major operands: CTA pairs win whenever there are enough 256-row tiles to occupy every SM pair; the tile
    //    width is the one that pads N least (256 on ties: fewer A re-reads).
    //  * MN-major operands (wgrad, split-K): single-CTA 128 x 256 tiles are fastest even when N is padded.
    //  * few tiles: single CTAs (twice as many work items); shrink BN until the SMs are covered.
    const bool mn_major = a->a_major || a->b_major;
    auto padded = [&](int n) { return cdiv(a->N, n) * n; };
    int auto_bn = 256;
    if (!mn_major || a->N <= 192) {
      for (int n : {192, 128}) if (padded(n) < padded(auto_bn)) auto_bn = n;
    }
    if (bn == 0) bn = auto_bn;
    if (sim_epi && bn == 192) bn = 256;                        // the similarity epilogues are instantiated for 128 / 256
    if (ctas == 0) {
      const int64_t pair_tiles = cdiv(a->M, 2 * kBM) * cdiv(a->N, bn);
      ctas = (eb == 2 && !mn_major && pair_ok(bn) && pair_tiles >= ctx->num_sms / 2) ? 2 : 1;
      if (ctas == 1 && a->tile_n == 0 && !can_split) {
        while (bn > 128 && cdiv(a->M, kBM) * cdiv(a->N, bn) < ctx->num_sms) bn -= 64;
        if (sim_epi && bn == 192) bn = 128;
      }
    }
  }

  // ---- split-K weight gradients (both operands MN-major, fp32 atomic output): CTA pairs on 256 x 384|512 tiles.
  // Such a tile moves (32 KB + BN*128 B) of operands per 256 x BN x 64 MACs — 1.8x fewer L2->SM bytes per flop than
  // 128 x 256 single-CTA tiles, which are pinned at the ~11 TB/s L2 fabric limit because no two units of a split-K
  // GEMM ever read the same bytes.  dW or dW^T is computed, whichever wastes fewer padded rows/columns.
  simseg_gemm_args sw;
  GemmSim sim_sw;
  int wide_bn = 0, d_trans = 0;
  const bool wide_ok = eb == 2 && a->a_major && a->b_major && can_split && a->bias == nullptr && a->tile_n == 0 &&
                       (a->reserved & (16 | 128)) == 0 && a->M % 64 == 0 && a->N % 64 == 0 && kb_total >= 64 &&
                       (reinterpret_cast<uintptr_t>(a->d) & 15) == 0;
  if (wide_ok) {
    auto cost = [&](int64_t m, int64_t n, int& bn_out) {
      double best = 1e300;
      for (int cand : {384, 512}) {
        const double c = static_cast<double>(cdiv(m, 256) * 256) * static_cast<double>(cdiv(n, cand) * cand);
        if (c < best) { best = c; bn_out = cand; }
      }
      return best;
    };
    int bn_n = 0, bn_t = 0;
    const double c_n = cost(a->M, a->N, bn_n), c_t = cost(a->N, a->M, bn_t);
    // only worth it when the padded problem stays within ~35 % of the real one
    const double real = static_cast<double>(a->M) * static_cast<double>(a->N);
    if ((c_t < c_n ? c_t : c_n) <= 1.35 * real) {
      if (c_t < c_n) {
        sw = *a;
        sw.a = a->b; sw.b = a->a; sw.M = a->N; sw.N = a->M; sw.lda = a->ldb; sw.ldb = a->lda;
        a = &sw;
        if (sim) { sim_sw = *sim; sim_sw.a_lo = sim->b_lo; sim_sw.b_lo = sim->a_lo; sim = &sim_sw; }   // the three products are symmetric
        d_trans = 1;
        wide_bn = bn_t;
      } else {
        wide_bn = bn_n;
      }
      bn = wide_bn;
      ctas = 2;
    }
  }

  GemmParams p{};
  if (sim) {
    p.split3 = 1; p.kb_seg = kb_seg;
    p.temperature = sim->temperature; p.row_offset = sim->row_offset; p.part = reinterpret_cast<float4*>(sim->part);
    p.part_ld = sim->part_ld; p.zt = sim->zt; p.lse = sim->lse; p.grad_scale = sim->grad_scale; p.dtemp = sim->dtemp;
    p.lo_off = sim->lo_off; p.best = sim->best; p.bestj = sim->bestj; p.lgid = sim->lgid; p.rgid = sim->rgid; p.rank = sim->rank;
    p.bestkey = sim->bestkey; p.tile_list = sim->tile_list; p.tile_count = sim->tile_count;
    if (sim_epi == kEpiNceFwd) SIMSEG_CHECK_ARG(sim->part_ld >= 2 * cdiv(a->N, bn), "gemm: partial buffer too narrow for %d-wide tiles", bn);
  }
  p.d_trans = d_trans;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.a_mn = a->a_major ? 1 : 0; p.b_mn = a->b_major ? 1 : 0;
  p.elem_bytes = eb;
  p.m_tiles = static_cast<int>(cdiv(a->M, kBM * ctas));
  p.n_tiles = static_cast<int>(cdiv(a->N, bn));
  p.kb_total = kb_total;
  const int units_max = ctx->num_sms / ctas;
  // split-K only for plain fp32 accumulation outputs with few output tiles (wgrad: K = all tokens).
  // The split count is chosen for wave efficiency: work items (tiles x splits) should fill whole waves of
  // persistent units; among equally efficient choices the smallest split count wins (fewer atomics).
  int splits = 1;
  const int tiles_mn = p.m_tiles * p.n_tiles;
  if (can_split && p.kb_total >= 32) {
    const int max_splits = p.kb_total / 8 < 64 ? p.kb_total / 8 : 64;
    double best_eff = 0.0;
    for (int s = 1; s <= max_splits; ++s) {
      const int kbs = static_cast<int>(cdiv(p.kb_total, s));
      const int s_eff = static_cast<int>(cdiv(p.kb_total, kbs));
      const int64_t items = static_cast<int64_t>(tiles_mn) * s_eff;
      const int64_t waves = cdiv(items, units_max);
      // time ~ waves * (k-blocks per item + fixed per-item cost of ~6 k-blocks for prologue/epilogue)
      const double eff = static_cast<double>(tiles_mn) * p.kb_total / (static_cast<double>(waves) * units_max * (kbs + 6.0));
      if (eff > best_eff * 1.02) { best_eff = eff; splits = s_eff; }
    }
  }
  p.kb_per_split = static_cast<int>(cdiv(p.kb_total, splits));
  p.splits = static_cast<int>(cdiv(p.kb_total, p.kb_per_split));
  p.d = a->d; p.ldd = a->ldd;
  p.out_bf16 = a->out_dtype == SIMSEG_BF16;
  p.atomic_out = (p.splits > 1 || a->accumulate || wide_bn) ? 1 : 0;     // wide tiles only have the reduction store path
  p.bias = a->bias; p.residual = a->residual; p.ld_res = a->ld_res; p.res_bf16 = a->res_dtype == SIMSEG_BF16;
  p.aux = a->aux; p.ld_aux = a->ld_aux; p.row_scale = a->row_scale; p.col_sum = a->col_sum;
  const int ob = p.out_bf16 ? 2 : 4;
  bool vec = (a->ldd * ob) % 16 == 0 && (reinterpret_cast<uintptr_t>(a->d) & 15) == 0;
  if (a->residual) vec = vec && (a->ld_res * (p.res_bf16 ? 2 : 4)) % 16 == 0 && (reinterpret_cast<uintptr_t>(a->residual) & 15) == 0;
  if (a->aux) vec = vec && (a->ld_aux * 2) % 16 == 0 && (reinterpret_cast<uintptr_t>(a->aux) & 15) == 0;
  if (a->bias) vec = vec && (reinterpret_cast<uintptr_t>(a->bias) & 15) == 0;
  p.vec_ok = vec ? 1 : 0;
  p.dbg = (a->reserved & 3) | ((a->reserved & 512) ? 4 : 0);          // 4 = late pre-activation prefetch (A/B)

  if ((p.splits > 1 || wide_bn) && !a->accumulate) {
    // partial sums are reduced with atomics: start from zero (D is stored [N, M] when d_trans)
    const int64_t rows = d_trans ? a->N : a->M, width = d_trans ? a->M : a->N;
    SIMSEG_CUDA(cudaMemset2DAsync(a->d, a->ldd * 4, 0, width * 4, rows, st));
  }

  CUtensorMap tm[5];
  memset(tm, 0, sizeof(tm));
  CUtensorMap& ta = tm[0];
  CUtensorMap& tb = tm[1];
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
  if (sim) tma_epi = false;                                  // tm[3] / tm[4] carry the lo operands
  p.tma_epi = tma_epi ? 1 : 0;
  p.aux2 = a->aux2;
  if (tma_epi) {
    const CUtensorMapSwizzle s64 = CU_TENSOR_MAP_SWIZZLE_64B;
    if ((rc = make_tmap_sw(&tm[2], a->d, 2, a->M, a->N, a->ldd, 32, 32, s64))) return rc;
    if (a->aux && (rc = make_tmap_sw(&tm[3], a->aux, 2, a->M, a->N, a->ld_aux, 32, 32, s64))) return rc;
    if (a->aux2 && (rc = make_tmap_sw(&tm[4], a->aux2, 2, a->M, a->N, a->ld_aux2, 32, 32, s64))) return rc;
  }
  // MN-major operands whose MN extent is a whole number of 128-byte atoms take ONE 3-D box per k-block
  // (atoms past the matrix edge are zero-filled by TMA); ragged extents keep one 2-D box per atom.
  const bool no3d = (a->reserved & 4) != 0;
  const int b_rows = bn / ctas;                                // B columns staged per CTA
  p.a_3d = (p.a_mn && a->M % mn_atom == 0 && !no3d) ? 1 : 0;
  p.b_3d = (p.b_mn && a->N % mn_atom == 0 && !no3d) ? 1 : 0;
  if (!p.a_mn) rc = make_tmap(&ta, a->a, eb, a->M, a->K, a->lda, k_elems, kBM);
  else if (p.a_3d) rc = make_tmap_mn3d(&ta, a->a, eb, a->K, a->M, a->lda, k_elems, kBM / mn_atom);
  else rc = make_tmap(&ta, a->a, eb, a->K, a->M, a->lda, mn_atom, k_elems);
  if (rc) return rc;
  if (!p.b_mn) rc = make_tmap(&tb, a->b, eb, a->N, a->K, a->ldb, k_elems, b_rows);
  else if (wide_bn) rc = make_tmap_mn3d(&tb, a->b, eb, a->K, a->N, a->ldb, k_elems, 1);
  else if (p.b_3d) rc = make_tmap_mn3d(&tb, a->b, eb, a->K, a->N, a->ldb, k_elems, b_rows / mn_atom);
  else rc = make_tmap(&tb, a->b, eb, a->K, a->N, a->ldb, mn_atom, k_elems);
  if (rc) return rc;
  if (sim) {                                                 // same geometry, lo halves
    if (!p.a_mn) rc = make_tmap(&tm[3], sim->a_lo, eb, a->M, a->K, a->lda, k_elems, kBM);
    else if (p.a_3d) rc = make_tmap_mn3d(&tm[3], sim->a_lo, eb, a->K, a->M, a->lda, k_elems, kBM / mn_atom);
    else rc = make_tmap(&tm[3], sim->a_lo, eb, a->K, a->M, a->lda, mn_atom, k_elems);
    if (rc) return rc;
    if (!p.b_mn) rc = make_tmap(&tm[4], sim->b_lo, eb, a->N, a->K, a->ldb, k_elems, b_rows);
    else if (wide_bn) rc = make_tmap_mn3d(&tm[4], sim->b_lo, eb, a->K, a->N, a->ldb, k_elems, 1);
    else if (p.b_3d) rc = make_tmap_mn3d(&tm[4], sim->b_lo, eb, a->K, a->N, a->ldb, k_elems, b_rows / mn_atom);
    else rc = make_tmap(&tm[4], sim->b_lo, eb, a->K, a->N, a->ldb, mn_atom, k_elems);
    if (rc) return rc;
  }

  // persistent grid: one unit (CTA or CTA pair) per SM (pair); with fused column sums keep every unit on ONE n-tile
  // (unit count a multiple of n_tiles) so the sums stay in registers until the unit is done
  const int total = p.m_tiles * p.n_tiles * p.splits;
  int units = total < units_max ? total : units_max;
  if (tma_epi && a->col_sum && p.n_tiles <= units) units = (units / p.n_tiles) * p.n_tiles;
  const int grid = units * ctas;

  if (sim_epi == kEpiNceFwd) return dispatch_sim<kEpiNceFwd>(ctx, bn, ctas, tm, p, grid, st);
  if (sim_epi == kEpiNceBwd) return dispatch_sim<kEpiNceBwd>(ctx, bn, ctas, tm, p, grid, st);
  if (sim_epi == kEpiRank) return dispatch_sim<kEpiRank>(ctx, bn, ctas, tm, p, grid, st);
  if (sim_epi == kEpiBest) return dispatch_sim<kEpiBest>(ctx, bn, ctas, tm, p, grid, st);
  if (wide_bn == 384) return launch_gemm<384, SIMSEG_EPI_NONE, 2>(ctx, tm, p, grid, st);
  if (wide_bn == 512) return launch_gemm<512, SIMSEG_EPI_NONE, 2>(ctx, tm, p, grid, st);
  // measured (profiles/r02_gemm_act16_ab.txt): dGELU gains 7-16 % on every shape of the step; the forward GELU gains 22 % at
  // K = 384 (ViT-S fc1: the epilogue, not the MMAs, paces the tile) and loses 3-10 % at K = 768, where the 8-warp epilogue
  // already hides under twice as many MMAs
  const bool act16 = tma_epi && (a->reserved & 64) == 0 && (a->epilogue == SIMSEG_EPI_DGELU || a->K <= 512 || (a->reserved & 256));
  if (ctas == 2) {
    switch (bn) {
      case 128: return dispatch_epi<128, 2>(ctx, a->epilogue, tm, p, grid, st, act16);
      case 192: return dispatch_epi<192, 2>(ctx, a->epilogue, tm, p, grid, st, act16);
      default: return dispatch_epi<256, 2>(ctx, a->epilogue, tm, p, grid, st, act16);
    }
  }
  switch (bn) {
    case 128: return dispatch_epi<128, 1>(ctx, a->epilogue, tm, p, grid, st, act16);
    case 192: return dispatch_epi<192, 1>(ctx, a->epilogue, tm, p, grid, st, act16);
    default: return dispatch_epi<256, 1>(ctx, a->epilogue, tm, p, grid, st, act16);
  }
}

}  // namespace simseg
