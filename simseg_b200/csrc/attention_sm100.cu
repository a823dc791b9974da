// tcgen05 attention backward for head_dim 64 and S <= 256 (ViT: 197 tokens; BERT: 25 / 77).
//
// One (batch, head) per work item, persistent CTAs (one per SM), 10 warps:
//   warp 0      TMA producer: Q, dO (two 128-row query tiles, resident for the whole item) and K, V (128-key tiles,
//               double buffered) land in 128B-swizzled smem through 4-D tensor maps {d, head, token, batch} — rows past S
//               are zero-filled by TMA, so no tile ever reads a neighbouring batch element.
//   warp 1      one elected thread issues every tcgen05.mma (all accumulators live in TMEM, 512 columns):
//                 S  = Q_qt K_kt^T   [128 q x 128 keys]  cols   0..127      dP = dO_qt V_kt^T         cols 128..255
//                 dV_kt += P^T dO_qt [128 keys x 64]     cols 256..319      dK_kt += dS^T Q_qt        cols 320..383
//                 dQ_qt += dS K_kt   [128 q x 64]        cols 384 + 64 qt
//   warps 2..9  elementwise, two phases per block (A: tcgen05.ld S, P = exp2(S*scale*log2e - lse*log2e); B: tcgen05.ld dP,
//               dS = P * (dP - D)) so the tensor pipe computes S of the NEXT block under phase B and dP of the next block under
//               its phase A (software-pipelined issue order, see the MMA warp); P and dS are written as bf16
//               into smem in the [key atom][q row][128 B] swizzled layout that serves BOTH as the K-major A operand of
//               dQ = dS K and as the MN-major A operand of dK = dS^T Q / dV = P^T dO (no transposes anywhere);
//               the same warps drain dK/dV after the last query tile of a key tile and dQ after the last key tile.
//   D_i = sum_d dO[i,d] O[i,d] is computed in-kernel from the dO tile in smem and one 128-byte global read of O per row.
// Loop order: key tile outer, query tile inner, so dK/dV accumulate in TMEM over the inner loop and dQ over the outer one;
// nothing is ever re-read from HBM and S x S never exists outside TMEM/smem.
#include "common.cuh"
#include "sm100.cuh"
#include "philox.cuh"

#include <cstdlib>

namespace simseg {

using namespace sm100;

// Drain of dK / dV (dQ) parked until after phase B of the NEXT block, with S(n) issued ahead of the accumulator wait: right after
// phase A the accumulators are still being written (the drain waited ~540 clk for dkv_full), a phase later they have landed.
// (A/B on the ViT-S layer: 2.23 -> 2.12 ms.)
// Single-block items (BERT, T = 77) keep the early drain: there every block ends a key tile, and the late drain would make the
// next block's dV wait for it (1.51 -> 1.75 ms).  The host picks the instantiation (nqt > 1).
constexpr int kAbEwWarps = 16;                       // elementwise warps: four per TMEM lane quarter
constexpr int kAbThreads = 32 * (2 + kAbEwWarps + 2);   // + two D / lse warps
constexpr int kTile = 128;                 // query rows / key rows per tile
constexpr int kTileBytes = kTile * 128;    // [128 rows][64 bf16]
constexpr int kPBytes = 2 * kTileBytes;    // P or dS: [2 key atoms][128 q rows][128 B]

struct AttnBwdParams {
  int32_t B, H, S, nqt, nkt, items;
  int32_t G, lg, HG, rows, tile_tx;   // packed mode (G > 1): G heads of one sequence per tile, row = token * G + head
  float scale, scale_log2e;
  const int32_t* key_len;
  const float* lse;                 // [B,H,S]
  const __nv_bfloat16* out;         // [B,S,H*64]
  const __nv_bfloat16* dout;        // [B,S,H*64]
  __nv_bfloat16 *dq, *dk, *dv;      // strides as q/k/v
  int64_t sb, ss, sh;
  int32_t dbg;                      // timing knock-outs (SIMSEG_ATTN_DBG, results wrong): see tools/attn_knockout.py
  const uint32_t* drop_mask;        // dropout keep bits, layout as in AttnFwdParams (kDrop instantiations only)
  int32_t mask_nw;
  float inv_keep;
};

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0, int32_t c1,
                                            int32_t c2, int32_t c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// 4-D tiled store shared -> global (bulk async group); elements outside the tensor are clipped
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem_src, int32_t c0, int32_t c1, int32_t c2, int32_t c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

// Timing knock-outs of the backward (SIMSEG_ATTN_DBG bits, results wrong: tools/attn_knockout.py) exist only in builds with
// -DSIMSEG_ATTN_KNOCKOUT: tested at run time they cost a predicate / select per element in the unrolled elementwise loops of the
// production kernel, whose instruction footprint is its second largest stall reason (no_inst, profiles/r02_ncu_attention.txt).
#ifdef SIMSEG_ATTN_KNOCKOUT
#define KO(bit) ((p.dbg & (bit)) != 0)
#else
#define KO(bit) false
#endif

__device__ __forceinline__ int ceil16(int x) { return (x + 15) & ~15; }

// 0 or 0xffffffff from bit `pos` of `word` (signed one-bit field extract)
__device__ __forceinline__ uint32_t bit_mask(uint32_t word, int pos) {
  int32_t r;
  asm("bfe.s32 %0, %1, %2, 1;" : "=r"(r) : "r"(static_cast<int32_t>(word)), "r"(pos));
  return static_cast<uint32_t>(r);
}
// mask of a packed bf16 pair from bits pos (low half) and pos + 1 (high half)
__device__ __forceinline__ uint32_t pair_mask(uint32_t word, int pos) {
  return (bit_mask(word, pos) & 0xffffu) | (bit_mask(word, pos + 1) & 0xffff0000u);
}
// Live key columns of a 32-column chunk as ONE word per thread instead of three compares per element: a column is live for a
// row iff it belongs to the row's head (packed tiles: column = token * G + head -> every G-th bit, `pat`) and lies below the
// caption's key length (`nlive` = live columns counted from the chunk's first column).
__device__ __forceinline__ uint32_t head_pattern(int G, int rg) {
  return (G == 1 ? 0xffffffffu : G == 2 ? 0x55555555u : G == 4 ? 0x11111111u : 0x01010101u) << rg;
}
__device__ __forceinline__ uint32_t live_mask(int nlive, uint32_t pat) {
  return nlive >= 32 ? pat : (nlive <= 0 ? 0u : (pat & ((1u << nlive) - 1u)));
}

// ---- debug: per-warp phase timing of CTA 0, read back with simseg_debug_trace_read (tools/attn_trace.py) ----------------
// Only in builds with -DSIMSEG_ATTN_TRACE (the register cost perturbs the kernel: use for RELATIVE shares only), armed with
// simseg_debug_trace_enable(1).  Slot 0 = MMA warp, slots 1.. = elementwise warps 2, 3, 6, 10.
constexpr int kTraceIds = 24;
constexpr int kTraceWarps = 5;
constexpr int kTraceLen = 2 * kTraceIds;                    // per slot: clocks per event id, then counts
__device__ unsigned long long g_attn_trace[kTraceWarps * kTraceLen];
__device__ int g_attn_trace_on = 0;
// Non-perturbing: clock deltas are summed in registers per event id ("time spent reaching this event from the previous
// one") and written out once at the end of the kernel — no memory traffic inside the loops.
#ifdef SIMSEG_ATTN_TRACE
struct Tracer {
  // one clock register per traced thread; deltas leave through fire-and-forget global reductions (RED, no return value, no
  // register arrays: the first version kept 48 accumulators per thread, spilled, and measured mostly its own reloads)
  uint32_t last;
  int slot;
  __device__ __forceinline__ Tracer(int slot_, bool active) : last(0), slot(-1) {
    if (active && g_attn_trace_on && blockIdx.x == 0 && slot_ >= 0) slot = slot_;
    last = static_cast<uint32_t>(clock64());
  }
  __device__ __forceinline__ void operator()(int id) {       // id must be a compile-time constant at every call site
    if (slot >= 0) {
      const uint32_t now = static_cast<uint32_t>(clock64());
      atomicAdd(&g_attn_trace[slot * kTraceLen + id], static_cast<unsigned long long>(now - last));
      atomicAdd(&g_attn_trace[slot * kTraceLen + kTraceIds + id], 1ull);
      last = static_cast<uint32_t>(clock64());
    }
  }
  __device__ __forceinline__ void flush() {}
};
#else
struct Tracer {                                              // compiled out: no code
  __device__ __forceinline__ Tracer(int, bool) {}
  __device__ __forceinline__ void operator()(int) {}
  __device__ __forceinline__ void flush() {}
};
#endif

// kDrop (dropout on the probabilities, BERT under model.train()): with keep bits m and c = 1 / (1 - p) the forward was
// O = c (m o P) V, so dV = c (m o P)^T dO (m o P goes to smem, c is applied in the dV drain), dP = c m o (dO V^T) and
// dS = P o (dP - D) with D = rowsum(dO o O) unchanged — P itself stays unmasked in registers between the two phases.
template <bool kLateDrain, bool kDrop>
__global__ void __launch_bounds__(kAbThreads, 1)
attention_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                        const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                        const __grid_constant__ CUtensorMap tm_dq, const __grid_constant__ CUtensorMap tm_dk,
                        const __grid_constant__ CUtensorMap tm_dv, const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
  uint8_t* sQ = smem;                          // [2][16 KB]
  uint8_t* sdO = sQ + 2 * kTileBytes;          // [2][16 KB]
  uint8_t* sK = sdO + 2 * kTileBytes;          // [2 buffers][16 KB]
  uint8_t* sV = sK + 2 * kTileBytes;           // [2 buffers][16 KB]
  uint8_t* sP = sV + 2 * kTileBytes;           // 32 KB
  uint8_t* sdS = sP + kPBytes;                 // 32 KB
  uint8_t* sStg = sdS + kPBytes;               // [16 warps][2 KB] drain staging (1024-byte aligned: SWIZZLE_64B TMA stores)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStg + kAbEwWarps * 2048);
  uint64_t* qdo_full = bars;          // [2]
  uint64_t* qdo_empty = bars + 2;     // [2]
  uint64_t* kv_full = bars + 4;       // [2]
  uint64_t* kv_empty = bars + 6;      // [2]
  uint64_t* s_full = bars + 8;        // MMA -> EW : S of block g is in TMEM
  uint64_t* dp_full = bars + 9;       // MMA -> EW : dP of block g is in TMEM
  uint64_t* p_ready = bars + 10;      // EW -> MMA : S consumed, P written to smem               (8 warps)
  uint64_t* ds_ready = bars + 11;     // EW -> MMA : dP consumed, dS written to smem             (8 warps)
  uint64_t* p_free = bars + 12;       // MMA -> EW : dV = P^T dO has read P
  uint64_t* ds_free = bars + 13;      // MMA -> EW : dK = dS^T Q and dQ = dS K have read dS
  uint64_t* dkv_full = bars + 14;     // MMA -> EW : dK/dV of a key tile complete
  uint64_t* dkv_free = bars + 15;     // EW -> MMA : dK/dV drained                               (8 warps)
  uint64_t* dq_full = bars + 16;
  uint64_t* dq_free = bars + 17;      //                                                          (8 warps)
  uint64_t* d_full = bars + 18;       // [2] producer -> EW : D = rowsum(dO * O) and lse * log2(e) of a Q / dO slot are in smem
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  // per Q / dO slot and tile row: {D, lse * log2 e}, written by the producer warp (it has the dO tile in smem and 31 idle
  // lanes) one item ahead of the elementwise warps — their only global loads and a 4-warp exchange used to sit here
  float2* sDL = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [2 slots][128 rows]
  if (threadIdx.x == 0 && static_cast<uint32_t>(smem - smem_raw) > 768u) {
    printf("simseg: attention_bwd dynamic shared memory base is not 256-byte aligned\n");
    __trap();
  }

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_q); prefetch_tmap(&tm_k); prefetch_tmap(&tm_v); prefetch_tmap(&tm_do);
    prefetch_tmap(&tm_dq); prefetch_tmap(&tm_dk); prefetch_tmap(&tm_dv);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qdo_full[i], 1); mbar_init(&qdo_empty[i], 1);
      mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1); mbar_init(dp_full, 1); mbar_init(p_ready, kAbEwWarps); mbar_init(ds_ready, kAbEwWarps);
    mbar_init(p_free, 1); mbar_init(ds_free, 1);
    mbar_init(dkv_full, 1); mbar_init(dkv_free, kAbEwWarps); mbar_init(dq_full, 1); mbar_init(dq_free, kAbEwWarps);
    mbar_init(&d_full[0], 1); mbar_init(&d_full[1], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (p.G > 1) {
    // packed boxes cover only S*G rows of a tile: the rows behind them must hold finite values (0 x NaN = NaN in the MMAs)
    uint4* z = reinterpret_cast<uint4*>(sQ);
    for (int i = threadIdx.x; i < 8 * kTileBytes / 16; i += kAbThreads) z[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tdP = tmem_base + 128, tdV = tmem_base + 256, tdK = tmem_base + 320, tdQ = tmem_base + 384;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      uint32_t kvc = 0, it = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        const int b = item / p.HG, h = (item - b * p.HG) * p.G;
        for (int kt = 0; kt < p.nkt; ++kt, ++kvc) {
          const uint32_t buf = kvc & 1, ph = (kvc >> 1) & 1;
          mbar_wait(&kv_empty[buf], ph ^ 1);
          mbar_arrive_expect_tx(&kv_full[buf], 2 * p.tile_tx);
          tma_load_4d(sK + buf * kTileBytes, &tm_k, &kv_full[buf], 0, h, kt * kTile, b);
          tma_load_4d(sV + buf * kTileBytes, &tm_v, &kv_full[buf], 0, h, kt * kTile, b);
          if (kt == 0) {
            // Q / dO slots: one per query tile; single-tile items alternate between the two slots, so the next item's
            // tiles load while the current item is still being worked on (the MMA issue runs one block ahead)
            for (int qt = 0; qt < p.nqt; ++qt) {
              const uint32_t slot = (p.nqt == 1) ? (it & 1) : static_cast<uint32_t>(qt);
              const uint32_t use = (p.nqt == 1) ? (it >> 1) : it;
              mbar_wait(&qdo_empty[slot], (use & 1) ^ 1);
              mbar_arrive_expect_tx(&qdo_full[slot], 2 * p.tile_tx);
              tma_load_4d(sQ + slot * kTileBytes, &tm_q, &qdo_full[slot], 0, h, qt * kTile, b);
              tma_load_4d(sdO + slot * kTileBytes, &tm_do, &qdo_full[slot], 0, h, qt * kTile, b);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    // Warp-uniform loops, one elected lane issues (values stay in uniform registers; a descriptor is a 64-bit add away
    // from two constants).  Issued from inside `if (lane == 0)` this thread — ~45 instructions per MMA — was the
    // slowest stage of the kernel.
    // Software pipeline over the flattened block sequence (item, key tile, query tile): the elementwise warps work in two
    // phases per block — A: S -> P, B: dP -> dS — and the tensor pipe is fed in the order
    //     [A(g) done] dV(g), S(g+1)      [B(g) done] dK(g), dQ(g), dP(g+1)
    // so S(g+1) is computed under phase B of block g and dP(g+1) under phase A of block g+1: the elementwise warps, not the
    // commit -> wait -> ld -> st -> fence -> arrive round trips, set the pace (round 1 ran the five products and the
    // elementwise pass of a block strictly one after the other: 70 % of the kernel was hand-off latency).
    {
      const uint32_t id_sdp_base = make_idesc(1u, 0u, 0u, kTile, 16u) & ~(0x3Fu << 17);     // N filled per key tile
      const uint32_t id_dkv = make_idesc(1u, 1u, 1u, kTile, 64u);                           // A, B MN-major
      const uint32_t id_dq = make_idesc(1u, 0u, 1u, kTile, 64u);                            // A K-major, B MN-major
      const uint64_t kd = make_smem_desc_sw128(0, 16, 1024);                                // K-major operand template
      const uint64_t md = make_smem_desc_sw128(0, 16384, 1024);                             // MN-major operand template
      const uint32_t aP = smem_u32(sP) >> 4, adS = smem_u32(sdS) >> 4;
      struct Cur { int item, kt, qt; uint32_t it, kvc, g; };
      auto valid = [&](const Cur& c) { return c.item < p.items; };
      auto advance = [&](Cur& c) {
        ++c.g;
        if (++c.qt == p.nqt) {
          c.qt = 0; ++c.kvc;
          if (++c.kt == p.nkt) { c.kt = 0; c.item += gridDim.x; ++c.it; }
        }
      };
      auto nkc_of = [&](const Cur& c) { return min(kTile, ceil16(p.rows - c.kt * kTile)); };
      auto slot_of = [&](const Cur& c) { return (p.nqt == 1) ? (c.it & 1) : static_cast<uint32_t>(c.qt); };
      auto use_of = [&](const Cur& c) { return (p.nqt == 1) ? (c.it >> 1) : c.it; };
      auto idesc_sdp = [&](const Cur& c) { return id_sdp_base | (static_cast<uint32_t>(nkc_of(c) >> 3) << 17); };
      auto issue_s = [&](const Cur& c) {              // S = Q K^T   (K = d = 64: four K-steps inside one 128-byte swizzle row)
        const uint32_t buf = c.kvc & 1;
        if (c.qt == 0) mbar_wait(&kv_full[buf], (c.kvc >> 1) & 1);
        if (c.kt == 0) mbar_wait(&qdo_full[slot_of(c)], use_of(c) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t aQ = smem_u32(sQ + slot_of(c) * kTileBytes) >> 4, aK = smem_u32(sK + buf * kTileBytes) >> 4;
          const uint32_t id = idesc_sdp(c);
#pragma unroll
          if (!KO(32))
            for (int kk = 0; kk < 4; ++kk) umma_f16(tS, kd + aQ + 2 * kk, kd + aK + 2 * kk, id, kk > 0 ? 1u : 0u);
          umma_commit(s_full);
        }
        __syncwarp();
      };
      auto issue_dp = [&](const Cur& c) {             // dP = dO V^T  (operands were waited for by issue_s of the same block)
        const uint32_t buf = c.kvc & 1;
        if (elect_one()) {
          const uint32_t adO = smem_u32(sdO + slot_of(c) * kTileBytes) >> 4, aV = smem_u32(sV + buf * kTileBytes) >> 4;
          const uint32_t id = idesc_sdp(c);
#pragma unroll
          if (!KO(32))
            for (int kk = 0; kk < 4; ++kk) umma_f16(tdP, kd + adO + 2 * kk, kd + aV + 2 * kk, id, kk > 0 ? 1u : 0u);
          umma_commit(dp_full);
        }
        __syncwarp();
      };
      uint32_t drains = 0;
      Tracer tr(0, lane == 0);
      Cur c{static_cast<int>(blockIdx.x), 0, 0, 0u, 0u, 0u}, n = c;
      if (valid(n)) { issue_s(n); issue_dp(n); advance(n); }
      while (valid(c)) {
        tr(1);
        const uint32_t buf = c.kvc & 1;
        const uint32_t slot = slot_of(c);
        const uint32_t aQ = smem_u32(sQ + slot * kTileBytes) >> 4, adO = smem_u32(sdO + slot * kTileBytes) >> 4;
        const uint32_t aK = smem_u32(sK + buf * kTileBytes) >> 4;
        const int qsteps = min(kTile, ceil16(p.rows - c.qt * kTile)) >> 4;   // K-steps over query rows
        const int ksteps = nkc_of(c) >> 4;                                   // K-steps over keys
        // ---- phase A of block c is done: P is in smem, the S buffer is free
        mbar_wait(p_ready, c.g & 1);
        tr(2);
        if (kLateDrain && valid(n)) { tc_fence_after(); issue_s(n); }           // S buffer is free: ahead of the accumulator wait
        if (c.qt == 0 && drains > 0) mbar_wait(dkv_free, (drains - 1) & 1);  // dV/dK accumulators drained (first MMA overwrites)
        tr(3);
        tc_fence_after();
        if (elect_one()) {
          // dV += P^T dO      (A MN-major: key atoms 16 KB apart; K-step = 16 q rows = 2048 B)
          for (int ks = 0; ks < (KO(16) ? 0 : qsteps); ++ks)
            umma_f16(tdV, md + aP + 128 * ks, md + adO + 128 * ks, id_dkv, (c.qt > 0 || ks > 0) ? 1u : 0u);
          umma_commit(p_free);
        }
        __syncwarp();
        tr(4);
        if (!kLateDrain && valid(n)) issue_s(n);                             // runs under phase B of block c
        tr(5);
        // ---- phase B of block c is done: dS is in smem, the dP buffer is free
        mbar_wait(ds_ready, c.g & 1);
        tr(6);
        if (c.kt == 0 && c.qt == 0 && c.it > 0) mbar_wait(dq_free, (c.it - 1) & 1);
        tc_fence_after();
        if (valid(n)) { issue_dp(n); advance(n); }                           // ahead of dK / dQ: phase B of the next block waits for it
        if (elect_one()) {
          // dK += dS^T Q
          for (int ks = 0; ks < (KO(16) ? 0 : qsteps); ++ks)
            umma_f16(tdK, md + adS + 128 * ks, md + aQ + 128 * ks, id_dkv, (c.qt > 0 || ks > 0) ? 1u : 0u);
          // dQ += dS K        (A K-major: 4 K-steps per 64-key atom; B MN-major: K-step = 16 key rows)
          for (int ks = 0; ks < (KO(16) ? 0 : ksteps); ++ks)
            umma_f16(tdQ + 64 * c.qt, kd + adS + (ks >> 2) * (kTileBytes >> 4) + (ks & 3) * 2, md + aK + 128 * ks, id_dq,
                     (c.kt > 0 || ks > 0) ? 1u : 0u);
          umma_commit(ds_free);
          if (c.qt == p.nqt - 1) {
            umma_commit(dkv_full);
            umma_commit(&kv_empty[buf]);
          }
          if (c.kt == p.nkt - 1) {
            umma_commit(&qdo_empty[slot]);
            if (c.qt == p.nqt - 1) umma_commit(dq_full);
          }
        }
        __syncwarp();
        tr(7);
        if (c.qt == p.nqt - 1) ++drains;
        advance(c);
      }
      tr.flush();
    }
  } else if (warp < 2 + kAbEwWarps) {
    // =============================== elementwise + drains ===============================
    // 16 warps: four per TMEM lane quarter (a warp may only touch lanes 32 * (warp % 4) ..), each owning ONE 32-key chunk of
    // a block.  The pass is latency-bound (dependent MUFU / FMA / pack chains, smem round trips), so it is spread over four
    // warps per scheduler instead of two; per block a thread handles 32 S and 32 dP values of its query row.
    const int ew = warp - 2;                   // 0..15
    const int quarter = warp & 3;              // TMEM lane quarter
    const int cq = ew >> 2;                    // 32-column chunk of a block handled by this warp
    const int r = quarter * 32 + lane;         // row inside a tile
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const int sw = r & 7;
    const int col0 = cq * 32;
    // drain staging: a thread owns a tile ROW (32 of its 64 d-columns = 64 bytes); stored straight from registers every
    // instruction would touch 32 different 128-byte lines.  The warp's 32 rows go through a SWIZZLE_64B 2 KB block and leave
    // as ONE 4-D TMA store {32 d, G heads, 32/G tokens, 1 batch}: no address arithmetic, rows past S are clipped by the
    // tensor map, and — unlike st.global — the stores are not waited for by the MEMBAR of the next fence.proxy.async.
    uint8_t* stg = sStg + ew * 2048;
    uint32_t it = 0, g = 0, drains = 0;
    Tracer tr(warp == 2 ? 1 : warp == 3 ? 2 : warp == 6 ? 3 : warp == 10 ? 4 : -1, lane == 0);
    // w[16] = this thread's row (32 bf16 of d-columns dcol0..dcol0+31) -> rows tile_row0 .. +32 of the tensor behind `tm`
    auto store_rows = [&](const uint32_t (&w)[16], const CUtensorMap* tm, int sb_idx, int sh0, int tile_row0, int dcol0) {
      if (lane == 0) tma_store_wait_read<0>();                  // the block's previous store has been read (long ago)
      __syncwarp();
      uint8_t* mine = stg + lane * 64;
      const int swz = (lane >> 1) & 3;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(mine + ((j ^ swz) << 4)) = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0 && tile_row0 < p.rows && !KO(8)) {
        tma_store_4d(tm, stg, dcol0, sh0, tile_row0 >> p.lg, sb_idx);
        tma_store_commit();
      }
    };
    // Deferred drains: the accumulators of a key tile (dK / dV) and of an item (dQ) complete ~a block's worth of MMAs after
    // this warp's last contribution; draining right away would idle the elementwise warps for that long (20 % of their
    // time).  The drain is parked and done after phase A of the NEXT block instead, when the accumulators have long landed.
    bool pend_kv = false, pend_q = false;
    int pend_b = 0, pend_h0 = 0, pend_kt = 0;
    uint32_t pend_it = 0;
    auto drain_pending = [&]() {
      const bool is_dk = cq >= 2;
      const int dcol0 = (cq & 1) * 32;
      if (pend_kv) {
        // dV (warps cq = 0, 1) and dK (warps cq = 2, 3) of key tile pend_kt: thread = key row, 32 of the 64 d-columns
        mbar_wait(dkv_full, drains & 1);
        tr(19);
        tc_fence_after();
        uint32_t a[32];
        tmem_ld_32x32((is_dk ? tdK : tdV) + lane_off + dcol0, a);
        tmem_ld_wait();
        // the accumulator is in registers: release it BEFORE the global stores (an mbarrier arrive has release
        // semantics — placed after the stores it would wait for them to drain)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dkv_free);
        const float sc = is_dk ? p.scale : (kDrop ? p.inv_keep : 1.0f);
        uint32_t w[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) w[j] = pack_bf16(__uint_as_float(a[2 * j]) * sc, __uint_as_float(a[2 * j + 1]) * sc);
        store_rows(w, is_dk ? &tm_dk : &tm_dv, pend_b, pend_h0, pend_kt * kTile + quarter * 32, dcol0);
        ++drains;
        pend_kv = false;
        tr(20);
      }
      if (pend_q) {
        // dQ: warp cq takes query tile cq >> 1, d-columns 32 (cq & 1) ..
        mbar_wait(dq_full, pend_it & 1);
        tr(21);
        tc_fence_after();
        const int dqt = cq >> 1;
        uint32_t w[16];
        if (dqt < p.nqt) {
          uint32_t a2[32];
          tmem_ld_32x32(tdQ + 64 * dqt + lane_off + dcol0, a2);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) w[j] = pack_bf16(__uint_as_float(a2[2 * j]) * p.scale, __uint_as_float(a2[2 * j + 1]) * p.scale);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dq_free);
        if (dqt < p.nqt) store_rows(w, &tm_dq, pend_b, pend_h0, dqt * kTile + quarter * 32, dcol0);
        pend_q = false;
        tr(22);
      }
    };
    // dropout keep word of (item, key tile, query tile) for this thread's row and 32-key chunk; all-ones where the chunk or the
    // row quarter is not live (never used there)
    auto load_keep = [&](int item_, int kt_, int qt_) -> uint32_t {
      const int nkc_ = min(kTile, ceil16(p.rows - kt_ * kTile));
      if (!(col0 < nkc_ && qt_ * kTile + quarter * 32 < ceil16(p.rows))) return 0xffffffffu;
      const int qrow_ = qt_ * kTile + r;
      return __ldg(p.drop_mask + (static_cast<int64_t>(item_) * p.rows + min(qrow_, p.rows - 1)) * p.mask_nw + kt_ * 4 + cq);
    };
    uint32_t keep_next = 0xffffffffu;
    if (kDrop && static_cast<int>(blockIdx.x) < p.items) keep_next = load_keep(blockIdx.x, 0, 0);
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      const int b = item / p.HG, h0 = (item - b * p.HG) * p.G;
      // klen = live key COLUMNS (packed: key tokens * G, column = token * G + head); a column is live for a row iff it
      // belongs to the row's head
      const int klen = (p.key_len ? min(max(p.key_len[b], 1), p.S) : p.S) * p.G;
      const int gm = p.G - 1, rg = r & gm;
      const bool need_mask = p.key_len != nullptr || p.G > 1;
      for (int kt = 0; kt < p.nkt; ++kt) {
        const int nkc = min(kTile, ceil16(p.rows - kt * kTile));
        const bool chunk_live = col0 < nkc;                           // warp-uniform: columns >= nkc are never read by an MMA
        for (int qt = 0; qt < p.nqt; ++qt, ++g) {
          const int qrow = qt * kTile + r;
          const bool q_ok = qrow < p.rows;
          // rows of this quarter that no MMA reads (K extent over query rows = ceil16(live rows)): skip the whole pass
          const bool rows_live = qt * kTile + quarter * 32 < ceil16(p.rows);
          const int qslot = (p.nqt == 1) ? static_cast<int>(it & 1) : qt;     // Q / dO slot (single-tile items alternate slots)
          const uint32_t quse = (p.nqt == 1) ? (it >> 1) : it;
          // ---- phase A: S -> P (kept as packed bf16 in registers for phase B), P into smem
          // keep bits of this thread's (row, 32-key chunk): the word was requested a whole block ago (keep_next), the word of
          // the NEXT block — same item or the CTA's next item — is requested now, so the global load is never waited for
          uint32_t keep = 0xffffffffu;
          if (kDrop) {
            keep = keep_next;
            int n_item = item, n_kt = kt, n_qt = qt + 1;
            if (n_qt == p.nqt) { n_qt = 0; if (++n_kt == p.nkt) { n_kt = 0; n_item += gridDim.x; } }
            if (n_item < p.items) keep_next = load_keep(n_item, n_kt, n_qt);
          }
          tr(10);
          mbar_wait(s_full, g & 1);                                    // implies the Q / dO tiles of this qt have landed
          tr(11);
          tc_fence_after();
          // D = rowsum(dO * O) and lse (log2 units) of this row: written by the producer warp, long before S is ready
          mbar_wait(&d_full[qslot], quse & 1);
          const float2 dl = sDL[qslot * kTile + r];
          const float Dq = dl.x, Lq = dl.y;
          tr(12);
          uint32_t pp[16];
          if (chunk_live && rows_live && !KO(128)) {
            uint32_t sr[32];
            if (!KO(64)) {
              tmem_ld_32x32(tS + lane_off + col0, sr);
              tmem_ld_wait();
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) sr[j] = 0;
            }
            if (need_mask) {                                       // padding keys: exp2(-lse) could overflow, mask them
              const uint32_t live = q_ok ? live_mask(klen - (kt * kTile + col0), head_pattern(p.G, rg)) : 0u;
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                const float p0 = ex2_approx(fmaf(__uint_as_float(sr[j]), p.scale_log2e, -Lq));
                const float p1 = ex2_approx(fmaf(__uint_as_float(sr[j + 1]), p.scale_log2e, -Lq));
                pp[j >> 1] = pack_bf16(p0, p1) & pair_mask(live, j);       // dead keys (possibly inf) -> exactly 0
              }
            } else if (kt * kTile + col0 + 32 > p.rows) {
              // last, partial chunk of an unmasked, unpacked sequence (ViT: key 197 of 208): only its live keys, warp-uniform
              // early exit; dead query rows as in the dense path below
              const int rem = p.rows - (kt * kTile + col0);
#pragma unroll
              for (int j = 0; j < 16; ++j) pp[j] = 0u;
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                if (j >= rem) break;
                const float p0 = ex2_approx(fmaf(__uint_as_float(sr[j]), p.scale_log2e, -Lq));
                const float p1 = j + 1 < rem ? ex2_approx(fmaf(__uint_as_float(sr[j + 1]), p.scale_log2e, -Lq)) : 0.f;
                pp[j >> 1] = pack_bf16(p0, p1);
              }
            } else {
              // dense, unmasked chunk of live keys: query rows past S come from zero-filled TMA rows (Q = dO = 0, lse := 0), so
              // P = 1, dS = 0 there and every product they enter is exactly zero — no per-element predicates needed
              const f32x2 sc2 = f2_splat(p.scale_log2e), nl2 = f2_splat(-Lq);
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                float e0, e1;
                f2_unpack(f2_fma(f2_pack(__uint_as_float(sr[j]), __uint_as_float(sr[j + 1])), sc2, nl2), e0, e1);
                pp[j >> 1] = KO(1) ? pack_bf16(e0, e1) : pack_bf16(ex2_approx(e0), ex2_approx(e1));
              }
            }
          }
          tr(13);
          if (g > 0) mbar_wait(p_free, (g - 1) & 1);                   // dV of the previous block has read sP
          tr(14);
          // 32 keys = four 16-byte chunks of this row inside key atom (col0 / 64)
          const uint32_t rowoff = (col0 >> 6) * kTileBytes + r * 128;
          const int ch0 = (col0 & 63) >> 3;
          if (chunk_live && rows_live && !KO(4)) {
            if (kDrop) {
              auto pm = [&](int i) {                                   // pair i = keys 2 i, 2 i + 1 of the chunk
                // bfe.s32 of one bit = 0 or ~0: two field extracts, one select-by-constant (LOP3), one AND — no predicates
                return pp[i] & ((bit_mask(keep, 2 * i) & 0xffffu) | (bit_mask(keep, 2 * i + 1) & 0xffff0000u));
              };
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4)
                *reinterpret_cast<uint4*>(sP + rowoff + (((ch0 + q4) ^ sw) << 4)) =
                    make_uint4(pm(4 * q4), pm(4 * q4 + 1), pm(4 * q4 + 2), pm(4 * q4 + 3));
            } else {
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4)
                *reinterpret_cast<uint4*>(sP + rowoff + (((ch0 + q4) ^ sw) << 4)) =
                    make_uint4(pp[4 * q4], pp[4 * q4 + 1], pp[4 * q4 + 2], pp[4 * q4 + 3]);
            }
          }
          tc_fence_before();
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(p_ready);
          tr(15);
          if (!kLateDrain) drain_pending();                              // accumulators of the previous key tile / item

          // ---- phase B: dP -> dS = P * (dP - D), dS into smem
          mbar_wait(dp_full, g & 1);
          tr(16);
          tc_fence_after();
          if (g > 0) mbar_wait(ds_free, (g - 1) & 1);
          tr(17);                  // dK / dQ of the previous block have read sdS (issued a
                                                                       // whole phase A ago: this wait does not stall)
          if (chunk_live && rows_live && !KO(2)) {
            uint32_t dr[32], dd[16];
            if (!KO(64)) {
              tmem_ld_32x32(tdP + lane_off + col0, dr);
              tmem_ld_wait();
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) dr[j] = 0;
            }
            const f32x2 nd2 = f2_splat(-Dq);
            const f32x2 ik2 = f2_splat(p.inv_keep);
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const uint32_t pk = pp[j >> 1];                          // masked entries are exactly 0 -> dS = 0
              float d0, d1;
              if (kDrop) {
                // dropped keys: dP := 0 by masking its bits, then one packed FMA with the splat 1 / (1 - p)
                f2_unpack(f2_mul(f2_pack(bf16_lo(pk), bf16_hi(pk)),
                                 f2_fma(f2_pack(__uint_as_float(dr[j] & bit_mask(keep, j)), __uint_as_float(dr[j + 1] & bit_mask(keep, j + 1))),
                                        ik2, nd2)), d0, d1);
              } else {
                f2_unpack(f2_mul(f2_pack(bf16_lo(pk), bf16_hi(pk)),
                                 f2_add(f2_pack(__uint_as_float(dr[j]), __uint_as_float(dr[j + 1])), nd2)), d0, d1);
              }
              dd[j >> 1] = pack_bf16(d0, d1);
            }
            if (!KO(4))
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4)
              *reinterpret_cast<uint4*>(sdS + rowoff + (((ch0 + q4) ^ sw) << 4)) =
                  make_uint4(dd[4 * q4], dd[4 * q4 + 1], dd[4 * q4 + 2], dd[4 * q4 + 3]);
          }
          tc_fence_before();
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(ds_ready);
          tr(18);
          if (kLateDrain) drain_pending();

          if (qt == p.nqt - 1) {                                      // park the drains (see drain_pending)
            pend_kv = true; pend_b = b; pend_h0 = h0; pend_kt = kt;
            if (kt == p.nkt - 1) { pend_q = true; pend_it = it; }
          }
        }
      }
    }
    drain_pending();
    if (lane == 0) tma_store_wait<0>();
    tr.flush();
  } else {
    // =============================== D / lse rows (warps 18, 19) ===============================
    // D_i = sum_d dO[i,d] O[i,d] and lse_i * log2(e) for the 128 rows of Q / dO slot `w` — one warp per slot, both operands
    // straight from HBM / L2 (8 lanes fetch one 128-byte row, so a load instruction covers four full lines; 16 of them in
    // flight), results parked in registers: the warp runs a whole item AHEAD of the elementwise warps and only has to drop its
    // 128 {D, lse} pairs into shared memory when the slot's previous use has been released.  (Computed by the elementwise
    // warps at the head of every query tile, these loads were the longest exposed latency of the kernel.)
    const int w = warp - 2 - kAbEwWarps;                        // slot served by this warp
    const int sub = lane >> 3, c8 = lane & 7;
    const int gm = p.G - 1;
    float2* dl = sDL + w * kTile;
    uint32_t it = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      if (p.nqt == 1 ? (static_cast<int>(it & 1) != w) : (w >= p.nqt)) continue;
      const int qt = (p.nqt == 1) ? 0 : w;
      const uint32_t use = (p.nqt == 1) ? (it >> 1) : it;
      const int b = item / p.HG, h = (item - b * p.HG) * p.G;
      const int nrows = min(kTile, p.rows - qt * kTile);
#ifndef SIMSEG_ATTN_NO_D_PREFETCH
      {
        // pull the O / dO rows of this warp's NEXT tile into L2 now: its loads (a whole item later) then see L2 latency, and so
        // does the TMA load of that dO tile
        const int nitem = item + static_cast<int>(gridDim.x) * (p.nqt == 1 ? 2 : 1);
        if (nitem < p.items) {
          const int nb = nitem / p.HG, nh = (nitem - nb * p.HG) * p.G;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = lane + 32 * i;
            if (r < nrows) {
              const int row = qt * kTile + r;
              const int64_t off = (static_cast<int64_t>(nb) * p.S + (row >> p.lg)) * (p.H * 64) + (nh + (row & gm)) * 64;
              asm volatile("prefetch.global.L2 [%0];" ::"l"(p.out + off));
              asm volatile("prefetch.global.L2 [%0];" ::"l"(p.dout + off));
            }
          }
        }
      }
#endif
      float dv[4], lv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {                              // this lane keeps rows (8 i + c8) * 4 + sub
        const int r = (8 * i + c8) * 4 + sub;
        lv[i] = 0.f;
        dv[i] = 0.f;
        if (r < nrows) {
          const int row = qt * kTile + r;
          lv[i] = __ldg(p.lse + (static_cast<int64_t>(b) * p.H + h + (row & gm)) * p.S + (row >> p.lg));
        }
      }
      // a ROLLED loop over the four 32-row groups: unrolled, this block was 3.1 k of the kernel's 7.2 k SASS instructions (32 rows x
      // 8 products + 96 shuffles), streamed once per item by two warps through the instruction cache the elementwise warps live in
#pragma unroll 1
      for (int i = 0; i < 4; ++i) {
        if (32 * i >= nrows) break;                              // warp-uniform
        uint4 o[8], a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int r = (8 * i + j) * 4 + sub;
          o[j] = make_uint4(0, 0, 0, 0);
          a[j] = make_uint4(0, 0, 0, 0);
          if (r < nrows) {
            const int row = qt * kTile + r;
            const int64_t off = (static_cast<int64_t>(b) * p.S + (row >> p.lg)) * (p.H * 64) + (h + (row & gm)) * 64;
            o[j] = __ldg(reinterpret_cast<const uint4*>(p.out + off) + c8);
            a[j] = __ldg(reinterpret_cast<const uint4*>(p.dout + off) + c8);
          }
        }
        float mine = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float part = bf16_lo(a[j].x) * bf16_lo(o[j].x) + bf16_hi(a[j].x) * bf16_hi(o[j].x) + bf16_lo(a[j].y) * bf16_lo(o[j].y) +
                       bf16_hi(a[j].y) * bf16_hi(o[j].y) + bf16_lo(a[j].z) * bf16_lo(o[j].z) + bf16_hi(a[j].z) * bf16_hi(o[j].z) +
                       bf16_lo(a[j].w) * bf16_lo(o[j].w) + bf16_hi(a[j].w) * bf16_hi(o[j].w);
          part += __shfl_xor_sync(0xffffffffu, part, 1);
          part += __shfl_xor_sync(0xffffffffu, part, 2);
          part += __shfl_xor_sync(0xffffffffu, part, 4);
          if (c8 == j) mine = part;
        }
        dv[0] = i == 0 ? mine : dv[0];                           // static indices: dv stays in registers
        dv[1] = i == 1 ? mine : dv[1];
        dv[2] = i == 2 ? mine : dv[2];
        dv[3] = i == 3 ? mine : dv[3];
      }
      mbar_wait(&qdo_empty[w], (use & 1) ^ 1);                  // the slot's previous use (and its readers of dl) is over
#pragma unroll
      for (int i = 0; i < 4; ++i) dl[(8 * i + c8) * 4 + sub] = make_float2(dv[i], lv[i] * 1.44269504088896341f);
      __syncwarp();
      if (lane == 0) mbar_arrive(&d_full[w]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ================================================================================================
// tcgen05 attention FORWARD for head_dim 64 and S <= 224 (ViT: 197 tokens; BERT: 77).
//
// Work unit = (batch, head, 128-row query tile).  All keys of a sequence fit one TMEM accumulator, so the softmax is
// exact two-pass (row max, then exp / sum) with no online rescaling:
//   warp 0      TMA: K, V of an item (all keys, double buffered across items), Q tile per unit (double buffered)
//   warp 1      one elected thread issues  S_u = Q_u K^T  (128 x keys, TMEM cols [0|stride) by unit parity) and, one
//               unit later, O_u = P_u V (128 x 64, TMEM cols 2*stride..) — so QK^T of the next unit runs on the tensor
//               pipe while the softmax warps work on the current one
//   warps 2-9   softmax: two warps per TMEM lane quarter take alternate 32-key chunks: tcgen05.ld S, max, exp2 (MUFU — the
//               real bound of this kernel), row sum, P (bf16) into 128B-swizzled smem as the K-major A operand of P V; the
//               same warps drain O of the previous unit, scale by 1/sum and store the output rows + the log-sum-exp.
// Keys >= key_len[b] (BERT padding) and the zero-filled tail columns get P = 0.
// Short sequences (S * G <= 128, e.g. BERT's 25 tokens): G heads of one sequence are PACKED into one tile by a single
// TMA box {64 d, G heads, S tokens} (row = token * G + head); S = Q K^T is then block-sparse and entries whose row and
// column belong to different heads are masked like padding — 4x fewer units, each with full-size MMAs.
constexpr int kAfThreads = 320;

struct AttnFwdParams {
  int32_t B, H, S, nqt, nkt, nkc, stride, items, kv_stages;
  int32_t nbuf;        // S accumulators in TMEM: 2 (alternating units) while 2 * stride + 64 <= 512, else 1 (long sequences)
  int32_t kv_tiles;    // 16 KB tiles per K (and per V) array = nkt * kv_stages
  int32_t p_atoms;     // 64-key atoms of the P tile in smem (tc variant) = ceil(nkc / 64)
  int32_t G, lg, HG, rows, tile_tx;   // packed mode (G > 1): G heads of one sequence share a 128-row tile, row = token * G + head
  float scale_log2e;
  const int32_t* key_len;
  float* lse;                 // [B,H,S]
  __nv_bfloat16* out;         // [B,S,H*64]
  // dropout on the probabilities (HF BertSelfAttention under model.train()): keep bits of tile row r = token * G + head,
  // 32 tile columns per word — mask[(item * rows + r) * mask_nw + chunk], written by attn_dropout_mask_kernel; kept P are
  // scaled by inv_keep = 1 / (1 - p) (folded into the output normalisation)
  const uint32_t* drop_mask;
  int32_t mask_nw;
  float inv_keep;
};

__global__ void __launch_bounds__(kAfThreads, 1)
attention_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                        const __grid_constant__ CUtensorMap tm_v, const AttnFwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;                    // every byte of the 227 KB is used: no alignment slack
  if ((smem_u32(smem_raw) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("simseg: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* sQ = smem;                          // [2 units][16 KB]
  uint8_t* sK = sQ + 2 * kTileBytes;           // [kv_stages items][nkt tiles x 16 KB]  rows = keys, contiguous over the tiles
  uint8_t* sV = sK + p.kv_tiles * kTileBytes;  // (2 stages of 2 tiles, 4 stages of 1 tile when S <= 128, 1 stage of 3 tiles when S > 256)
  uint8_t* sP = sV + p.kv_tiles * kTileBytes;  // [p_atoms key atoms][128 rows][128 B]
  float* s_xchg = reinterpret_cast<float*>(sP + p.p_atoms * kTileBytes);     // [256 row sums | 256 row maxima] exchanged between the two warps of a row
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_xchg + 512);
  uint64_t* q_full = bars;            // [2]
  uint64_t* q_empty = bars + 2;       // [2]
  uint64_t* kv_full = bars + 4;       // [4]
  uint64_t* kv_empty = bars + 8;      // [4]
  uint64_t* s_full = bars + 12;       // [2] MMA -> softmax group g
  uint64_t* s_empty = bars + 14;      // [2] group g -> MMA (4 warps)
  uint64_t* p_ready = bars + 16;      // softmax -> MMA (4 warps)
  uint64_t* p_free = bars + 17;       // MMA -> softmax : P V has read P
  uint64_t* o_full = bars + 18;       // MMA -> softmax group : O complete
  uint64_t* o_free = bars + 19;       // group -> MMA (4 warps) : O drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_q); prefetch_tmap(&tm_k); prefetch_tmap(&tm_v);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1);
      mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 8);
    }
    for (int i = 0; i < 4; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    mbar_init(p_ready, 8); mbar_init(p_free, 1); mbar_init(o_full, 1); mbar_init(o_free, 8);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (p.G > 1) {
    // packed boxes cover only S*G rows of a tile: the rows behind them must hold finite values (P = 0 times NaN is NaN)
    uint4* z = reinterpret_cast<uint4*>(sQ);
    for (int i = threadIdx.x; i < 10 * kTileBytes / 16; i += kAfThreads) z[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tO = tmem_base + p.nbuf * p.stride;
  const uint32_t bmask = static_cast<uint32_t>(p.nbuf - 1);        // S buffer of unit u = u & bmask, its phase = (u / nbuf) & 1
  const uint32_t bshift = bmask;                                    // nbuf = 2 -> 1, nbuf = 1 -> 0

  if (warp == 0) {
    // =============================== TMA producer ===============================
    uint32_t u = 0, itc = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++itc) {
      const int b = item / p.HG, h = (item - b * p.HG) * p.G;
      const uint32_t kb = itc % p.kv_stages;
      mbar_wait(&kv_empty[kb], ((itc / p.kv_stages) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&kv_full[kb], 2 * p.nkt * p.tile_tx);
        for (int t = 0; t < p.nkt; ++t) {
          tma_load_4d(sK + (p.nkt * kb + t) * kTileBytes, &tm_k, &kv_full[kb], 0, h, t * kTile, b);
          tma_load_4d(sV + (p.nkt * kb + t) * kTileBytes, &tm_v, &kv_full[kb], 0, h, t * kTile, b);
        }
      }
      __syncwarp();
      for (int qt = 0; qt < p.nqt; ++qt, ++u) {
        mbar_wait(&q_empty[u & 1], ((u >> 1) & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&q_full[u & 1], p.tile_tx);
          tma_load_4d(sQ + (u & 1) * kTileBytes, &tm_q, &q_full[u & 1], 0, h, qt * kTile, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    // A, B K-major, N = keys; one instruction covers at most 256 keys, longer sequences (S = 325 at 288 x 288) take a second one
    const uint32_t n0 = static_cast<uint32_t>(p.nkc > 256 ? 256 : p.nkc), n1 = static_cast<uint32_t>(p.nkc) - n0;
    const uint32_t id_s = make_idesc(1u, 0u, 0u, kTile, n0);
    const uint32_t id_s1 = make_idesc(1u, 0u, 0u, kTile, n1 ? n1 : 16u);
    const uint32_t id_o = make_idesc(1u, 0u, 1u, kTile, 64u);                            // A K-major (P), B MN-major (V)
    const uint32_t aP = smem_u32(sP);
    const int ksteps = p.nkc >> 4;
    uint32_t u = 0, itc = 0;
    uint32_t prev_kb = 0, prev_last = 0;
    auto issue_pv = [&](uint32_t w, uint32_t kb, uint32_t last) {
      mbar_wait(p_ready, w & 1);
      mbar_wait(o_free, (w & 1) ^ 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t aV = smem_u32(sV + p.nkt * kb * kTileBytes);
        for (int ks = 0; ks < ksteps; ++ks)
          umma_f16(tO, make_smem_desc_sw128(aP + (ks >> 2) * kTileBytes + (ks & 3) * 32, 16, 1024),
                   make_smem_desc_sw128(aV + ks * 2048, 16384, 1024), id_o, ks > 0 ? 1u : 0u);
        umma_commit(o_full);
        umma_commit(p_free);
        if (last) umma_commit(&kv_empty[kb]);
      }
      __syncwarp();
    };
    bool pv_owed = false;                        // P V of unit u-1 not issued yet
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++itc) {
      const uint32_t kb = itc % p.kv_stages;
      for (int qt = 0; qt < p.nqt; ++qt, ++u) {
        // single-stage K/V (long sequences): the next item's keys can only load once the LAST P V of this item has released
        // the buffer, so that P V must be issued before waiting for them (it is normally deferred behind the next S = Q K^T)
        if (qt == 0 && p.kv_stages == 1 && pv_owed) { issue_pv(u - 1, prev_kb, prev_last); pv_owed = false; }
        if (qt == 0) mbar_wait(&kv_full[kb], (itc / p.kv_stages) & 1);
        mbar_wait(&q_full[u & 1], (u >> 1) & 1);
        mbar_wait(&s_empty[u & bmask], ((u >> bshift) & 1) ^ 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t aQ = smem_u32(sQ + (u & 1) * kTileBytes), aK = smem_u32(sK + p.nkt * kb * kTileBytes);
          const uint32_t tS = tmem_base + (u & bmask) * p.stride;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_f16(tS, make_smem_desc_sw128(aQ + kk * 32, 16, 1024), make_smem_desc_sw128(aK + kk * 32, 16, 1024), id_s,
                     kk > 0 ? 1u : 0u);
          if (n1) {                                  // keys 256.. : K rows two tiles further, accumulator columns 256..
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_f16(tS + 256, make_smem_desc_sw128(aQ + kk * 32, 16, 1024),
                       make_smem_desc_sw128(aK + 2 * kTileBytes + kk * 32, 16, 1024), id_s1, kk > 0 ? 1u : 0u);
          }
          umma_commit(&s_full[u & bmask]);
          umma_commit(&q_empty[u & 1]);            // Q is only read by these MMAs: reload its buffer as early as possible
        }
        __syncwarp();
        if (pv_owed) issue_pv(u - 1, prev_kb, prev_last);
        pv_owed = true;
        prev_kb = kb;
        prev_last = (qt == p.nqt - 1) ? 1u : 0u;
      }
    }
    if (pv_owed) issue_pv(u - 1, prev_kb, prev_last);
  } else {
    // =============================== softmax + output (8 warps on every unit) ===============================
    // Two warps share a TMEM lane quarter (32 query rows) and take alternate 32-key chunks; the row max and the row
    // sum are combined through shared memory (one 64-thread named barrier each).  Order per unit u:
    //   pass 1 (max) of u | drain O of u-1 (its P V has long finished) | pass 2 (exp, P) of u
    // so the only wait on the tensor pipe is the short P V of u-1 before P is overwritten.
    const int half = (warp - 2) >> 2;          // chunk parity taken by this warp
    const int quarter = warp & 3;              // TMEM lane quarter
    const int r = quarter * 32 + lane;         // row inside the tile
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const int sw = r & 7;
    const int me = (half * 4 + quarter) * 32 + lane, mate = ((half ^ 1) * 4 + quarter) * 32 + lane;
    uint32_t u = 0;
    // state of the previous unit, kept for its deferred output
    bool pv_pending = false, p_q_ok = false, p_live = false;
    float p_sum = 1.f, p_ms = 0.f;
    int64_t p_row = 0, p_lse = 0;
    auto drain = [&](uint32_t w) {
      mbar_wait(o_full, w & 1);
      tc_fence_after();
      if (p_live) {
        uint32_t o0[32];
        tmem_ld_32x32(tO + lane_off + half * 32, o0);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_free);                            // O is in registers: release it before the stores
        if (p_q_ok) {
          const float inv = 1.0f / p_sum;
          __nv_bfloat16* po = p.out + p_row + half * 32;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 a;
            a.x = pack_bf16(__uint_as_float(o0[8 * j]) * inv, __uint_as_float(o0[8 * j + 1]) * inv);
            a.y = pack_bf16(__uint_as_float(o0[8 * j + 2]) * inv, __uint_as_float(o0[8 * j + 3]) * inv);
            a.z = pack_bf16(__uint_as_float(o0[8 * j + 4]) * inv, __uint_as_float(o0[8 * j + 5]) * inv);
            a.w = pack_bf16(__uint_as_float(o0[8 * j + 6]) * inv, __uint_as_float(o0[8 * j + 7]) * inv);
            reinterpret_cast<uint4*>(po)[j] = a;
          }
          if (half == 0 && p.lse) p.lse[p_lse] = (p_ms + log2f(p_sum)) * 0.69314718055994531f;
        }
      }
      if (!p_live) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_free);
      }
    };
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const int b = item / p.HG, h0 = (item - b * p.HG) * p.G;
      // klen = live key COLUMNS of the tile (packed: key tokens * G, column = token * G + head)
      const int klen = (p.key_len ? min(max(p.key_len[b], 1), p.S) : p.S) * p.G;
      const int nch = (klen + 31) >> 5;                                    // 32-column chunks that hold a live key
      const int gm = p.G - 1, rg = r & gm;                                 // a column is live iff it is this row's head
      const bool dense = p.G == 1;
      for (int qt = 0; qt < p.nqt; ++qt, ++u) {
        const int qrow = qt * kTile + r;                                    // packed: row = token * G + head
        const bool q_ok = qrow < p.rows;
        const bool warp_live = qt * kTile + quarter * 32 < p.rows;          // any live row in this warp
        const uint32_t tS = tmem_base + (u & bmask) * p.stride + lane_off;
        mbar_wait(&s_full[u & bmask], (u >> bshift) & 1);
        tc_fence_after();
        // ---- pass 1: row max over this warp's chunks (chunk = half + 2k), exchanged with the mate warp
        float m = -INFINITY;
        if (warp_live) {
          for (int c = half; c < nch; c += 2) {
            uint32_t x[32];
            tmem_ld_32x32(tS + c * 32, x);
            tmem_ld_wait();
            if (dense && c * 32 + 32 <= klen) {                            // whole chunk live: no per-key predicate
              float m0 = fmaxf(__uint_as_float(x[0]), __uint_as_float(x[1]));
              float m1 = fmaxf(__uint_as_float(x[2]), __uint_as_float(x[3]));
#pragma unroll
              for (int j = 4; j < 32; j += 2) {
                m0 = fmaxf(m0, __uint_as_float(x[j]));
                m1 = fmaxf(m1, __uint_as_float(x[j + 1]));
              }
              m = fmaxf(m, fmaxf(m0, m1));
            } else if (dense) {
              const int nv = klen - c * 32;                                // live columns of this (last) chunk
#pragma unroll
              for (int j = 0; j < 32; ++j) m = (j < nv) ? fmaxf(m, __uint_as_float(x[j])) : m;
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                m = (c * 32 + j < klen && ((c * 32 + j) & gm) == rg) ? fmaxf(m, __uint_as_float(x[j])) : m;
            }
          }
        }
        s_xchg[256 + me] = m;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
        m = fmaxf(m, s_xchg[256 + mate]);
        const float ms = m * p.scale_log2e;
        // ---- output of the previous unit (needs o_free before this unit's P V can start)
        if (pv_pending) drain(u - 1);
        // ---- pass 2: P = exp2(s*scale*log2e - max), row sums, bf16 P into the swizzled A-operand layout
        mbar_wait(p_free, (u & 1) ^ 1);                                   // P V of the previous unit has read sP
        float sum = 0.f;
        if (warp_live) {
          for (int c = half; c < nch; c += 2) {
            uint32_t x[32];
            tmem_ld_32x32(tS + c * 32, x);
            tmem_ld_wait();
            uint32_t pp[16];
            if (dense && c * 32 + 32 <= klen) {
              float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float p0 = ex2_approx(fmaf(__uint_as_float(x[j]), p.scale_log2e, -ms));
                const float p1 = ex2_approx(fmaf(__uint_as_float(x[j + 1]), p.scale_log2e, -ms));
                const float p2 = ex2_approx(fmaf(__uint_as_float(x[j + 2]), p.scale_log2e, -ms));
                const float p3 = ex2_approx(fmaf(__uint_as_float(x[j + 3]), p.scale_log2e, -ms));
                s0 += p0; s1 += p1; s2 += p2; s3 += p3;
                pp[j >> 1] = pack_bf16(p0, p1);
                pp[(j >> 1) + 1] = pack_bf16(p2, p3);
              }
              sum += (s0 + s1) + (s2 + s3);
            } else if (dense) {
              const int nv = klen - c * 32;
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                float p0 = ex2_approx(fmaf(__uint_as_float(x[j]), p.scale_log2e, -ms));
                float p1 = ex2_approx(fmaf(__uint_as_float(x[j + 1]), p.scale_log2e, -ms));
                p0 = (j < nv) ? p0 : 0.f;
                p1 = (j + 1 < nv) ? p1 : 0.f;
                sum += p0 + p1;
                pp[j >> 1] = pack_bf16(p0, p1);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                float p0 = ex2_approx(fmaf(__uint_as_float(x[j]), p.scale_log2e, -ms));
                float p1 = ex2_approx(fmaf(__uint_as_float(x[j + 1]), p.scale_log2e, -ms));
                p0 = (c * 32 + j < klen && ((c * 32 + j) & gm) == rg) ? p0 : 0.f;
                p1 = (c * 32 + j + 1 < klen && ((c * 32 + j + 1) & gm) == rg) ? p1 : 0.f;
                sum += p0 + p1;
                pp[j >> 1] = pack_bf16(p0, p1);
              }
            }
            const uint32_t rowoff = (c >> 1) * kTileBytes + r * 128;
            const int ch0 = (c & 1) * 4;
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4)
              *reinterpret_cast<uint4*>(sP + rowoff + (((ch0 + q4) ^ sw) << 4)) =
                  make_uint4(pp[4 * q4], pp[4 * q4 + 1], pp[4 * q4 + 2], pp[4 * q4 + 3]);
          }
          // keys between the last live chunk and the MMA's K extent (nkc) must read as P = 0
          for (int c = nch + ((nch & 1) != half ? 1 : 0); c * 32 < p.nkc; c += 2) {
            const uint32_t rowoff = (c >> 1) * kTileBytes + r * 128;
            const int ch0 = (c & 1) * 4;
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4)
              *reinterpret_cast<uint4*>(sP + rowoff + (((ch0 + q4) ^ sw) << 4)) = make_uint4(0, 0, 0, 0);
          }
        }
        // row sum of the pair through shared memory (second barrier: the slots are rewritten by the next unit)
        s_xchg[me] = sum;
        tc_fence_before();
        fence_proxy_async_smem();
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
        sum += s_xchg[mate];
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
        if (lane == 0) { mbar_arrive(p_ready); mbar_arrive(&s_empty[u & bmask]); }
        pv_pending = true; p_q_ok = q_ok; p_live = warp_live; p_sum = sum; p_ms = ms;
        {
          const int tok = qrow >> p.lg, hh = h0 + (qrow & gm);
          p_row = (static_cast<int64_t>(b) * p.S + tok) * (p.H * 64) + hh * 64;
          p_lse = (static_cast<int64_t>(b) * p.H + hh) * p.S + tok;
        }
      }
    }
    if (pv_pending) drain(u - 1);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ================================================================================================
// Forward, variant "ts": P never leaves tensor memory.  S = Q K^T sits in one of two TMEM buffers; the softmax warps
// overwrite the head of the SAME buffer with P as packed bf16 pairs (tcgen05.st) and O = P V is issued with the A
// operand read from TMEM (tcgen05.mma [d], [a_tmem], b_desc).  No shared-memory P tile, no proxy fence, and — because P
// is double-buffered for free — two groups of four warps (one thread = one query row, no row-pair exchange) work on
// alternate units completely independently; only the MUFU rate couples them.
//   warp 0      TMA (as above)          warp 1   MMA issue: QK^T of unit u+1, then P V of unit u
//   warps 2-5   softmax + output of even units      warps 6-9   the same for odd units
// kDrop: the probabilities are dropped (keep bits from p.drop_mask) AFTER the row sum — softmax, then dropout, as
// BertSelfAttention does; log-sum-exp stays that of the undropped softmax.
template <bool kDrop>
__global__ void __launch_bounds__(kAfThreads, 1)
attention_fwd_ts_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                        const __grid_constant__ CUtensorMap tm_v, const AttnFwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                          // [2 units][16 KB]
  uint8_t* sK = sQ + 2 * kTileBytes;           // [kv_stages items][nkt tiles x 16 KB]
  uint8_t* sV = sK + p.kv_tiles * kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + p.kv_tiles * kTileBytes);
  uint64_t* q_full = bars;            // [2]
  uint64_t* q_empty = bars + 2;       // [2]
  uint64_t* kv_full = bars + 4;       // [4]
  uint64_t* kv_empty = bars + 8;      // [4]
  uint64_t* s_full = bars + 12;       // [2] MMA -> group g : S of a unit is complete
  uint64_t* p_ready = bars + 14;      // [2] group g -> MMA : P is in TMEM (4 warps)
  uint64_t* pv_done = bars + 16;      // [2] MMA -> MMA     : P V has read buffer g, it may be overwritten by the next S
  uint64_t* o_full = bars + 18;       // [2] MMA -> group g : O of one of ITS units is complete (per group: a group's first
                                      //     wait must be a parity-0 wait, or it would pass before anything completed)
  uint64_t* o_free = bars + 20;       // group -> MMA : O drained (4 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_q); prefetch_tmap(&tm_k); prefetch_tmap(&tm_v);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1);
      mbar_init(&s_full[i], 1); mbar_init(&p_ready[i], 4); mbar_init(&pv_done[i], 1);
    }
    for (int i = 0; i < 4; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    mbar_init(&o_full[0], 1); mbar_init(&o_full[1], 1); mbar_init(o_free, 4);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (p.G > 1) {
    uint4* z = reinterpret_cast<uint4*>(sQ);
    for (int i = threadIdx.x; i < 10 * kTileBytes / 16; i += kAfThreads) z[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tO = tmem_base + 2 * p.stride;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    uint32_t u = 0, itc = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++itc) {
      const int b = item / p.HG, h = (item - b * p.HG) * p.G;
      const uint32_t kb = itc % p.kv_stages;
      mbar_wait(&kv_empty[kb], ((itc / p.kv_stages) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&kv_full[kb], 2 * p.nkt * p.tile_tx);
        for (int t = 0; t < p.nkt; ++t) {
          tma_load_4d(sK + (p.nkt * kb + t) * kTileBytes, &tm_k, &kv_full[kb], 0, h, t * kTile, b);
          tma_load_4d(sV + (p.nkt * kb + t) * kTileBytes, &tm_v, &kv_full[kb], 0, h, t * kTile, b);
        }
      }
      __syncwarp();
      for (int qt = 0; qt < p.nqt; ++qt, ++u) {
        mbar_wait(&q_empty[u & 1], ((u >> 1) & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&q_full[u & 1], p.tile_tx);
          tma_load_4d(sQ + (u & 1) * kTileBytes, &tm_q, &q_full[u & 1], 0, h, qt * kTile, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    const uint32_t id_s = make_idesc(1u, 0u, 0u, kTile, static_cast<uint32_t>(p.nkc));   // A, B K-major, N = keys
    const uint32_t id_o = make_idesc(1u, 0u, 1u, kTile, 64u);                            // A (TMEM) K-major, B MN-major (V)
    const uint64_t kd = make_smem_desc_sw128(0, 16, 1024);
    const uint64_t md = make_smem_desc_sw128(0, 16384, 1024);
    const int ksteps = p.nkc >> 4;
    struct Cur { int item, qt; uint32_t itc, u; };
    auto valid = [&](const Cur& c) { return c.item < p.items; };
    auto advance = [&](Cur& c) {
      ++c.u;
      if (++c.qt == p.nqt) { c.qt = 0; c.item += gridDim.x; ++c.itc; }
    };
    auto issue_qk = [&](const Cur& c) {
      const uint32_t kb = c.itc % p.kv_stages, g = c.u & 1;
      if (c.qt == 0) mbar_wait(&kv_full[kb], (c.itc / p.kv_stages) & 1);
      mbar_wait(&q_full[g], (c.u >> 1) & 1);
      mbar_wait(&pv_done[g], ((c.u >> 1) & 1) ^ 1);          // P V of unit u-2 has read this buffer (passes for u < 2)
      tc_fence_after();
      if (elect_one()) {
        const uint32_t aQ = smem_u32(sQ + g * kTileBytes) >> 4, aK = smem_u32(sK + p.nkt * kb * kTileBytes) >> 4;
        const uint32_t tS = tmem_base + g * p.stride;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_f16(tS, kd + aQ + 2 * kk, kd + aK + 2 * kk, id_s, kk > 0 ? 1u : 0u);
        umma_commit(&s_full[g]);
        umma_commit(&q_empty[g]);
      }
      __syncwarp();
    };
    auto issue_pv = [&](const Cur& c) {
      const uint32_t kb = c.itc % p.kv_stages, g = c.u & 1;
      mbar_wait(&p_ready[g], (c.u >> 1) & 1);
      mbar_wait(o_free, (c.u & 1) ^ 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t aV = smem_u32(sV + p.nkt * kb * kTileBytes) >> 4;
        const uint32_t tP = tmem_base + g * p.stride;                        // bf16 pairs: 8 columns per 16-key k-step
        for (int ks = 0; ks < ksteps; ++ks) umma_f16_ts(tO, tP + 8 * ks, md + aV + 128 * ks, id_o, ks > 0 ? 1u : 0u);
        umma_commit(&o_full[g]);
        umma_commit(&pv_done[g]);
        if (c.qt == p.nqt - 1) umma_commit(&kv_empty[kb]);
      }
      __syncwarp();
    };
    Cur cq{static_cast<int>(blockIdx.x), 0, 0u, 0u}, cp = cq;
    if (valid(cq)) { issue_qk(cq); advance(cq); }
    while (valid(cp)) {
      if (valid(cq)) { issue_qk(cq); advance(cq); }
      issue_pv(cp);
      advance(cp);
    }
  } else {
    // =============================== softmax + output: group g owns units u = g (mod 2) ===============================
    const int g = (warp - 2) >> 2;
    const int quarter = warp & 3;              // TMEM lane quarter
    const int r = quarter * 32 + lane;         // row inside the tile
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    uint8_t* stg = reinterpret_cast<uint8_t*>(bars) + 256 + (warp - 2) * 4096;       // this warp's output staging block
    uint32_t u = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const int b = item / p.HG, h0 = (item - b * p.HG) * p.G;
      const int klen = (p.key_len ? min(max(p.key_len[b], 1), p.S) : p.S) * p.G;      // live key columns
      const int nch = (klen + 31) >> 5;
      const int gm = p.G - 1, rg = r & gm;
      const bool dense = p.G == 1;
      for (int qt = 0; qt < p.nqt; ++qt, ++u) {
        if ((u & 1) != static_cast<uint32_t>(g)) continue;
        const int qrow = qt * kTile + r;
        const bool q_ok = qrow < p.rows;
        const bool warp_live = qt * kTile + quarter * 32 < p.rows;
        const uint32_t tS = tmem_base + g * p.stride + lane_off;
        mbar_wait(&s_full[g], (u >> 1) & 1);
        tc_fence_after();
        float m = -INFINITY, sum = 0.f;
        const uint32_t* mrow = nullptr;
        uint32_t kw0 = 0xffffffffu, kw1 = 0xffffffffu, kw2 = 0xffffffffu, kw3 = 0xffffffffu;
        if (kDrop) {
          // the row's first four keep words (128 key columns: every packed case and T <= 128) are requested before pass 1 and
          // used in pass 2; longer rows read the remaining words in place
          mrow = p.drop_mask + (static_cast<int64_t>(item) * p.rows + (q_ok ? qrow : 0)) * p.mask_nw;
          if (warp_live) {
            kw0 = __ldg(mrow);
            if (nch > 1) kw1 = __ldg(mrow + 1);
            if (nch > 2) kw2 = __ldg(mrow + 2);
            if (nch > 3) kw3 = __ldg(mrow + 3);
          }
        }
        if (warp_live) {
          // ---- pass 1: row max
          for (int c = 0; c < nch; ++c) {
            uint32_t x[32];
            tmem_ld_32x32(tS + c * 32, x);
            tmem_ld_wait();
            if (dense && c * 32 + 32 <= klen) {
              // three-input max (FMNMX3): 16 instructions per 32 keys on two independent chains
              float m0 = fmax3(m, __uint_as_float(x[0]), __uint_as_float(x[1]));
              float m1 = fmaxf(__uint_as_float(x[2]), __uint_as_float(x[3]));
#pragma unroll
              for (int j = 4; j < 32; j += 4) {
                m0 = fmax3(m0, __uint_as_float(x[j]), __uint_as_float(x[j + 1]));
                m1 = fmax3(m1, __uint_as_float(x[j + 2]), __uint_as_float(x[j + 3]));
              }
              m = fmaxf(m0, m1);
            } else if (dense) {
              // last, partial chunk of an unpacked sequence (ViT: 197 = 6 x 32 + 5): only the live keys, warp-uniform early exit
              // (the per-element predicates of the general path cost more than the whole dense chunk for these 5 keys)
              const int rem = klen - c * 32;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (j >= rem) break;
                m = fmaxf(m, __uint_as_float(x[j]));
              }
            } else {
              const uint32_t live = live_mask(klen - c * 32, head_pattern(p.G, rg));
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const uint32_t bm = bit_mask(live, j);                        // dead keys enter the max as -inf
                m = fmaxf(m, __uint_as_float((x[j] & bm) | (0xff800000u & ~bm)));
              }
            }
          }
          const float ms = m * p.scale_log2e;
          // ---- pass 2: P = exp2(s*scale*log2e - max) as bf16 pairs into the first columns of the same TMEM buffer.
          //      Chunk c lands in columns [16c, 16c+16): always inside S columns this thread has already consumed.
          for (int c = 0; c * 32 < p.nkc; ++c) {
            uint32_t pp[16];
            if (c < nch) {
              uint32_t x[32];
              uint32_t keep = 0xffffffffu;
              if (kDrop) keep = c == 0 ? kw0 : c == 1 ? kw1 : c == 2 ? kw2 : c == 3 ? kw3 : __ldg(mrow + c);
              tmem_ld_32x32(tS + c * 32, x);
              tmem_ld_wait();
              if (dense && c * 32 + 32 <= klen) {
                // packed f32x2 for the scale / subtract and the row sums: 10 instead of 14 issue slots per four keys
                const f32x2 sc2 = f2_splat(p.scale_log2e), nm2 = f2_splat(-ms);
                f32x2 acc0 = f2_splat(0.f), acc1 = f2_splat(0.f);
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  float e0, e1, e2, e3;
                  f2_unpack(f2_fma(f2_pack(__uint_as_float(x[j]), __uint_as_float(x[j + 1])), sc2, nm2), e0, e1);
                  f2_unpack(f2_fma(f2_pack(__uint_as_float(x[j + 2]), __uint_as_float(x[j + 3])), sc2, nm2), e2, e3);
                  const float p0 = ex2_approx(e0), p1 = ex2_approx(e1), p2 = ex2_approx(e2), p3 = ex2_approx(e3);
                  acc0 = f2_add(acc0, f2_pack(p0, p1));
                  acc1 = f2_add(acc1, f2_pack(p2, p3));
                  if (kDrop) {
                    pp[j >> 1] = pack_bf16(p0, p1) & ((bit_mask(keep, j) & 0xffffu) | (bit_mask(keep, j + 1) & 0xffff0000u));
                    pp[(j >> 1) + 1] = pack_bf16(p2, p3) & ((bit_mask(keep, j + 2) & 0xffffu) | (bit_mask(keep, j + 3) & 0xffff0000u));
                  } else {
                    pp[j >> 1] = pack_bf16(p0, p1);
                    pp[(j >> 1) + 1] = pack_bf16(p2, p3);
                  }
                }
                float s0, s1, s2, s3;
                f2_unpack(acc0, s0, s1);
                f2_unpack(acc1, s2, s3);
                sum += (s0 + s1) + (s2 + s3);
              } else if (dense && !kDrop) {
                const int rem = klen - c * 32;
#pragma unroll
                for (int j = 0; j < 16; ++j) pp[j] = 0u;
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                  if (j >= rem) break;
                  const float p0 = ex2_approx(fmaf(__uint_as_float(x[j]), p.scale_log2e, -ms));
                  const float p1 = j + 1 < rem ? ex2_approx(fmaf(__uint_as_float(x[j + 1]), p.scale_log2e, -ms)) : 0.f;
                  sum += p0 + p1;
                  pp[j >> 1] = pack_bf16(p0, p1);
                }
              } else {
                const uint32_t live = live_mask(klen - c * 32, head_pattern(p.G, rg));
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                  float p0 = ex2_approx(fmaf(__uint_as_float(x[j]), p.scale_log2e, -ms));
                  float p1 = ex2_approx(fmaf(__uint_as_float(x[j + 1]), p.scale_log2e, -ms));
                  p0 = __uint_as_float(__float_as_uint(p0) & bit_mask(live, j));
                  p1 = __uint_as_float(__float_as_uint(p1) & bit_mask(live, j + 1));
                  sum += p0 + p1;
                  pp[j >> 1] = pack_bf16(p0, p1);
                  if (kDrop) pp[j >> 1] &= pair_mask(keep, j);
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) pp[j] = 0u;             // keys between the last live chunk and the MMA's K extent
            }
            tmem_st_32x16(tS + c * 16, pp);
          }
          tmem_st_wait();
          m = ms;                                                  // keep the scaled max for the log-sum-exp
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[g]);
        // ---- output: O / sum, log-sum-exp
        mbar_wait(&o_full[g], (u >> 1) & 1);
        tc_fence_after();
        if (warp_live) {
          uint32_t o0[32], o1[32];
          tmem_ld_32x32(tO + lane_off, o0);
          tmem_ld_32x32(tO + lane_off + 32, o1);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(o_free);                      // O is in registers: release it before the stores
          {
            // a thread owns a query ROW: stored straight from registers, every instruction would touch 32 different lines
            // (16 bytes each).  The warp's 32 rows x 128 bytes go through a swizzled 4 KB block and leave as 4 full
            // 128-byte rows per instruction.
            const float inv = q_ok ? (kDrop ? p.inv_keep / sum : 1.0f / sum) : 0.f;
            uint8_t* mine = stg + lane * 128;
            const int swz = lane & 7;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 a, c2;
              a.x = pack_bf16(__uint_as_float(o0[8 * j]) * inv, __uint_as_float(o0[8 * j + 1]) * inv);
              a.y = pack_bf16(__uint_as_float(o0[8 * j + 2]) * inv, __uint_as_float(o0[8 * j + 3]) * inv);
              a.z = pack_bf16(__uint_as_float(o0[8 * j + 4]) * inv, __uint_as_float(o0[8 * j + 5]) * inv);
              a.w = pack_bf16(__uint_as_float(o0[8 * j + 6]) * inv, __uint_as_float(o0[8 * j + 7]) * inv);
              c2.x = pack_bf16(__uint_as_float(o1[8 * j]) * inv, __uint_as_float(o1[8 * j + 1]) * inv);
              c2.y = pack_bf16(__uint_as_float(o1[8 * j + 2]) * inv, __uint_as_float(o1[8 * j + 3]) * inv);
              c2.z = pack_bf16(__uint_as_float(o1[8 * j + 4]) * inv, __uint_as_float(o1[8 * j + 5]) * inv);
              c2.w = pack_bf16(__uint_as_float(o1[8 * j + 6]) * inv, __uint_as_float(o1[8 * j + 7]) * inv);
              *reinterpret_cast<uint4*>(mine + ((j ^ swz) << 4)) = a;
              *reinterpret_cast<uint4*>(mine + (((j + 4) ^ swz) << 4)) = c2;
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = (lane >> 3) + 4 * i;
              const uint4 val = *reinterpret_cast<const uint4*>(stg + rr * 128 + (((lane & 7) ^ (rr & 7)) << 4));
              const int row = qt * kTile + quarter * 32 + rr;
              if (row < p.rows) {
                const int tok2 = row >> p.lg, hh2 = h0 + (row & gm);
                *reinterpret_cast<uint4*>(p.out + (static_cast<int64_t>(b) * p.S + tok2) * (p.H * 64) + hh2 * 64 + (lane & 7) * 8) = val;
              }
            }
            __syncwarp();
            if (q_ok && p.lse) {
              const int tok = qrow >> p.lg, hh = h0 + (qrow & gm);
              p.lse[(static_cast<int64_t>(b) * p.H + hh) * p.S + tok] = (m + log2f(sum)) * 0.69314718055994531f;
            }
          }
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(o_free);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn4)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// {64 d, H heads, S tokens, B batch} bf16 view of a (batch, token, head)-strided tensor; box = 128 tokens of one head
// G > 1: one box = {64 d, G heads, all S tokens} of one batch element (row = token * G + head in shared memory)
// store = true: box {32 d, G heads, 32 / G tokens, 1 batch}, SWIZZLE_64B — one elementwise warp's 32 rows x 64 bytes of dQ / dK / dV
static int make_tmap_bshd(CUtensorMap* m, const void* ptr, int B, int H, int S, int64_t sb, int64_t ss, int64_t sh, int G = 1,
                          bool store = false) {
  static EncodeTiledFn4 enc = nullptr;
  if (!enc) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      set_error("cuTensorMapEncodeTiled not available");
      return SIMSEG_ERR_CUDA;
    }
    enc = reinterpret_cast<EncodeTiledFn4>(f);
  }
  cuuint64_t dims[4] = {64, static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(S), static_cast<cuuint64_t>(B)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(sh) * 2, static_cast<cuuint64_t>(ss) * 2, static_cast<cuuint64_t>(sb) * 2};
  cuuint32_t box[4] = {64, static_cast<cuuint32_t>(G), static_cast<cuuint32_t>(G > 1 ? S : kTile), 1};
  if (store) { box[0] = 32; box[2] = static_cast<cuuint32_t>(32 / G); }
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, store ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(4d) failed (%d): ptr=%p B=%d H=%d S=%d sb=%lld ss=%lld sh=%lld", static_cast<int>(r), ptr, B, H, S,
              static_cast<long long>(sb), static_cast<long long>(ss), static_cast<long long>(sh));
    return SIMSEG_ERR_CUDA;
  }
  return SIMSEG_OK;
}

int debug_trace_enable_impl(int on) {
  SIMSEG_CUDA(cudaMemcpyToSymbol(g_attn_trace_on, &on, sizeof(int)));
  if (on) {
    void* ptr = nullptr;
    SIMSEG_CUDA(cudaGetSymbolAddress(&ptr, g_attn_trace));
    SIMSEG_CUDA(cudaMemset(ptr, 0, sizeof(unsigned long long) * kTraceWarps * kTraceLen));
  }
  return SIMSEG_OK;
}
int debug_trace_read_impl(unsigned long long* host, int n) {
  const int total = kTraceWarps * kTraceLen;
  SIMSEG_CUDA(cudaDeviceSynchronize());
  SIMSEG_CUDA(cudaMemcpyFromSymbol(host, g_attn_trace, sizeof(unsigned long long) * (n < total ? n : total)));
  return n < total ? n : total;
}

// heads of one sequence packed per 128-row tile: largest power of two G with G * S <= 128 that divides H
static int pack_factor(int H, int S) {
  int G = 1;
  while (G * 2 * S <= kTile && H % (G * 2) == 0 && G < 8) G *= 2;
  return G;
}

// ---- dropout keep bits for the probabilities, in the tile coordinates the two kernels above walk ---------------------------------
// Logical definition (oracle/simseg_oracle.py:attn_keep_mask): keep(b, h, q, k) = word (k & 3) of
// philox4x32_10(key = seed, counter = {k >> 2, (b H + h) S + q, site, step}) >= thr.  One thread per tile row r = q * G + head
// writes that row's ceil(rows / 32) words: bit j of word w = tile column 32 w + j = key token (32 w + j) >> lg of the head
// (32 w + j) & (G - 1); bits of other heads' columns and of columns >= rows are never read and stay 0.
__global__ void __launch_bounds__(256) attn_dropout_mask_kernel(uint32_t* __restrict__ mask, int B, int H, int S, int G, int lg,
                                                                const DropSpec ds) {
  const int rows = S * G, nw = (rows + 31) >> 5, HG = H / G;
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= static_cast<int64_t>(B) * HG * rows) return;
  const int r = static_cast<int>(t % rows);
  const int item = static_cast<int>(t / rows);
  const int b = item / HG, h = (item - b * HG) * G + (r & (G - 1)), q = r >> lg;
  const DropKey dk = load_drop_key(ds);
  const uint32_t row = static_cast<uint32_t>((b * H + h) * S + q);
  uint32_t* out = mask + t * nw;
  const int kpw = 32 >> lg;                                   // key tokens per word
  for (int w = 0; w < nw; ++w) {
    uint32_t bits = 0;
    for (int k0 = w * kpw; k0 < (w + 1) * kpw && k0 < S; k0 += 4) {       // kpw is 32, 16, 8 or 4: groups of four stay aligned
      const uint4 x = drop_words(ds, dk, static_cast<uint32_t>(k0 >> 2), row);
      const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = k0 + i;
        if (k < S && xs[i] >= ds.thr) bits |= 1u << ((k * G + (r & (G - 1))) & 31);
      }
    }
    out[w] = bits;
  }
}

int64_t attn_dropout_mask_words_impl(int B, int H, int S) {
  if (B < 1 || H < 1 || S < 1 || S > 256) return 0;
  const int G = pack_factor(H, S);
  const int rows = S * G;
  return static_cast<int64_t>(B) * (H / G) * rows * ((rows + 31) / 32);
}

int attn_dropout_mask_impl(Ctx* ctx, int B, int H, int S, float drop_p, const void* rng, uint32_t site, uint32_t* mask,
                           int64_t mask_words, cudaStream_t st) {
  const int64_t need = attn_dropout_mask_words_impl(B, H, S);
  SIMSEG_CHECK_ARG(need > 0 && mask_words >= need, "attn_dropout_mask: B=%d H=%d S=%d needs %lld words (S <= 256), buffer has %lld", B, H, S,
                   static_cast<long long>(need), static_cast<long long>(mask_words));
  SIMSEG_CHECK_ARG(static_cast<int64_t>(B) * H * S < (int64_t(1) << 32), "attn_dropout_mask: B*H*S exceeds the 32-bit row counter");
  const int G = pack_factor(H, S);
  const int lg = G == 1 ? 0 : (G == 2 ? 1 : (G == 4 ? 2 : 3));
  const int64_t threads = static_cast<int64_t>(B) * (H / G) * S * G;
  attn_dropout_mask_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, st>>>(mask, B, H, S, G, lg,
                                                                                        make_drop_spec(drop_p, rng, site));
  ctx->launches++;
  SIMSEG_LAUNCH_CHECK();
  return SIMSEG_OK;
}

// SIMSEG_ERR_UNSUPPORTED => caller uses the mma.sync kernel (S > 256, unaligned pointers, H == 1 with odd strides ...)
int attention_bwd_tc_impl(Ctx* ctx, const void* q, const void* k, const void* v, const void* out, const void* dout,
                          const float* lse, int64_t sb, int64_t ss, int64_t sh, int B, int H, int S, const int32_t* key_len,
                          float scale, void* dq, void* dk, void* dv, const uint32_t* drop_mask, float drop_p, cudaStream_t st) {
  if (S > 256 || S < 1) return SIMSEG_ERR_UNSUPPORTED;
  const uintptr_t al = reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
                       reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(dq) |
                       reinterpret_cast<uintptr_t>(dk) | reinterpret_cast<uintptr_t>(dv);
  if ((al & 15) || sb % 8 || ss % 8 || sh % 8) return SIMSEG_ERR_UNSUPPORTED;
  if ((H > 1 && sh * 2 >= (int64_t(1) << 40)) || ss * 2 >= (int64_t(1) << 40)) return SIMSEG_ERR_UNSUPPORTED;
  CUtensorMap tq, tk, tv, tdo;
  int rc;
  const int G = pack_factor(H, S);
  if ((rc = make_tmap_bshd(&tq, q, B, H, S, sb, ss, sh, G))) return rc;
  if ((rc = make_tmap_bshd(&tk, k, B, H, S, sb, ss, sh, G))) return rc;
  if ((rc = make_tmap_bshd(&tv, v, B, H, S, sb, ss, sh, G))) return rc;
  if ((rc = make_tmap_bshd(&tdo, dout, B, H, S, static_cast<int64_t>(S) * H * 64, static_cast<int64_t>(H) * 64, 64, G))) return rc;
  CUtensorMap tdq, tdk, tdv;
  if ((rc = make_tmap_bshd(&tdq, dq, B, H, S, sb, ss, sh, G, true))) return rc;
  if ((rc = make_tmap_bshd(&tdk, dk, B, H, S, sb, ss, sh, G, true))) return rc;
  if ((rc = make_tmap_bshd(&tdv, dv, B, H, S, sb, ss, sh, G, true))) return rc;
  AttnBwdParams p{};
  p.B = B; p.H = H; p.S = S;
  p.G = G; p.HG = H / G; p.rows = S * G;
  p.lg = G == 1 ? 0 : (G == 2 ? 1 : (G == 4 ? 2 : 3));
  p.tile_tx = G > 1 ? p.rows * 128 : kTileBytes;
  p.nqt = (p.rows + kTile - 1) / kTile; p.nkt = p.nqt;
  p.items = B * p.HG;
  p.scale = scale; p.scale_log2e = scale * 1.44269504088896341f;
  p.key_len = key_len; p.lse = lse; p.out = reinterpret_cast<const __nv_bfloat16*>(out); p.dout = reinterpret_cast<const __nv_bfloat16*>(dout);
  p.dq = reinterpret_cast<__nv_bfloat16*>(dq); p.dk = reinterpret_cast<__nv_bfloat16*>(dk); p.dv = reinterpret_cast<__nv_bfloat16*>(dv);
  p.sb = sb; p.ss = ss; p.sh = sh;
  { const char* d = getenv("SIMSEG_ATTN_DBG"); p.dbg = d ? atoi(d) : 0; }
  p.drop_mask = drop_mask; p.mask_nw = (p.rows + 31) / 32; p.inv_keep = drop_mask ? 1.0f / (1.0f - drop_p) : 1.0f;
  // alignment slack (768: the kernel traps if the base needs more) + tiles + P / dS + barriers + per-warp drain staging + {D, lse} rows
  const int smem_bytes = 768 + 8 * kTileBytes + 2 * kPBytes + 256 + kAbEwWarps * 2048 + 2 * kTile * 8;
  static bool attr_set = false;
  if (!attr_set) {
    SIMSEG_CUDA(cudaFuncSetAttribute(attention_bwd_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    SIMSEG_CUDA(cudaFuncSetAttribute(attention_bwd_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    SIMSEG_CUDA(cudaFuncSetAttribute(attention_bwd_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    SIMSEG_CUDA(cudaFuncSetAttribute(attention_bwd_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr_set = true;
  }
  const int grid = p.items < ctx->num_sms ? p.items : ctx->num_sms;
  if (drop_mask) {
    if (p.nqt > 1) attention_bwd_tc_kernel<true, true><<<grid, kAbThreads, smem_bytes, st>>>(tq, tk, tv, tdo, tdq, tdk, tdv, p);
    else attention_bwd_tc_kernel<false, true><<<grid, kAbThreads, smem_bytes, st>>>(tq, tk, tv, tdo, tdq, tdk, tdv, p);
  } else if (p.nqt > 1) attention_bwd_tc_kernel<true, false><<<grid, kAbThreads, smem_bytes, st>>>(tq, tk, tv, tdo, tdq, tdk, tdv, p);
  else attention_bwd_tc_kernel<false, false><<<grid, kAbThreads, smem_bytes, st>>>(tq, tk, tv, tdo, tdq, tdk, tdv, p);
  ctx->launches++;
  SIMSEG_LAUNCH_CHECK();
  return SIMSEG_OK;
}

// SIMSEG_ERR_UNSUPPORTED => caller uses the mma.sync kernel (S > 384, unaligned pointers ...)
// S <= 224: two S accumulators in TMEM, "ts" variant by default.  224 < S <= 384 (the reference's seg evaluation runs 288 x 288
// images = 325 tokens, configs/clip/simseg.vit-s.yaml:70-77): the shared-memory-P variant with ONE S accumulator (up to 384
// columns + 64 for O), all keys of a sequence in smem (3 K + 3 V tiles, single stage) and two MMAs per S = Q K^T (N <= 256 each).
int attention_fwd_tc_impl(Ctx* ctx, const void* q, const void* k, const void* v, int64_t sb, int64_t ss, int64_t sh, int B, int H,
                          int S, const int32_t* key_len, float scale, void* out, float* lse, const uint32_t* drop_mask, float drop_p,
                          cudaStream_t st) {
  if (S > 384 || S < 1) return SIMSEG_ERR_UNSUPPORTED;
  const int G = pack_factor(H, S);
  if (G == 1 && S < 48 && getenv("SIMSEG_ATTN_FWD") == nullptr && !drop_mask) return SIMSEG_ERR_UNSUPPORTED;   // mostly padding: mma.sync kernel
  const uintptr_t al = reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
                       reinterpret_cast<uintptr_t>(out);
  if ((al & 15) || sb % 8 || ss % 8 || sh % 8) return SIMSEG_ERR_UNSUPPORTED;
  if ((H > 1 && sh * 2 >= (int64_t(1) << 40)) || ss * 2 >= (int64_t(1) << 40)) return SIMSEG_ERR_UNSUPPORTED;
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_tmap_bshd(&tq, q, B, H, S, sb, ss, sh, G))) return rc;
  if ((rc = make_tmap_bshd(&tk, k, B, H, S, sb, ss, sh, G))) return rc;
  if ((rc = make_tmap_bshd(&tv, v, B, H, S, sb, ss, sh, G))) return rc;
  AttnFwdParams p{};
  p.B = B; p.H = H; p.S = S;
  p.G = G; p.HG = H / G; p.rows = S * G;
  p.lg = G == 1 ? 0 : (G == 2 ? 1 : (G == 4 ? 2 : 3));
  p.tile_tx = G > 1 ? p.rows * 128 : kTileBytes;
  p.nqt = (p.rows + kTile - 1) / kTile; p.nkt = p.nqt;
  p.nkc = (p.rows + 15) & ~15;
  p.stride = (p.nkc + 31) & ~31;
  p.items = B * p.HG;
  p.kv_stages = p.nkt == 1 ? 4 : (p.nkt == 2 ? 2 : 1);
  p.kv_tiles = p.nkt * p.kv_stages;
  p.nbuf = (2 * p.stride + 64 <= 512) ? 2 : 1;
  p.p_atoms = (p.nkc + 63) / 64 < 4 ? 4 : (p.nkc + 63) / 64;
  p.scale_log2e = scale * 1.44269504088896341f;
  p.key_len = key_len; p.lse = lse; p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.drop_mask = drop_mask; p.mask_nw = (p.rows + 31) / 32; p.inv_keep = drop_mask ? 1.0f / (1.0f - drop_p) : 1.0f;
  if (drop_mask && p.nbuf != 2) return SIMSEG_ERR_UNSUPPORTED;             // dropout lives in the "ts" variant only (S <= 224)
  // default: the "ts" variant (P stays in tensor memory; measured 0.895 vs 0.974 ms on 4096 x 6 x 197 and 0.174 vs 0.214 ms on
  // the packed 4096 x 12 x 25 case); SIMSEG_ATTN_FWD=tc selects the shared-memory-P variant below
  const char* var = getenv("SIMSEG_ATTN_FWD");
  if (p.nbuf == 2 && (drop_mask || !(var != nullptr && var[0] == 't' && var[1] == 'c'))) {
    const int smem_ts = 1024 + 10 * kTileBytes + 256 + 8 * 4096;                  // + per-warp output staging
    static bool ts_set = false;
    if (!ts_set) {
      SIMSEG_CUDA(cudaFuncSetAttribute(attention_fwd_ts_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_ts));
      SIMSEG_CUDA(cudaFuncSetAttribute(attention_fwd_ts_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_ts));
      ts_set = true;
    }
    const int grid_ts = p.items < ctx->num_sms ? p.items : ctx->num_sms;
    if (drop_mask) attention_fwd_ts_kernel<true><<<grid_ts, kAfThreads, smem_ts, st>>>(tq, tk, tv, p);
    else attention_fwd_ts_kernel<false><<<grid_ts, kAfThreads, smem_ts, st>>>(tq, tk, tv, p);
    ctx->launches++;
    SIMSEG_LAUNCH_CHECK();
    return SIMSEG_OK;
  }
  const int smem_bytes = (2 + 2 * p.kv_tiles + p.p_atoms) * kTileBytes + 2048 + 256;   // the dynamic segment itself is 1024-byte aligned (checked in-kernel)
  if (smem_bytes > 14 * kTileBytes + 2048 + 256) return SIMSEG_ERR_UNSUPPORTED;
  static bool attr_set = false;
  if (!attr_set) {
    SIMSEG_CUDA(cudaFuncSetAttribute(attention_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 14 * kTileBytes + 2048 + 256));
    attr_set = true;
  }
  const int grid = p.items < ctx->num_sms ? p.items : ctx->num_sms;
  attention_fwd_tc_kernel<<<grid, kAfThreads, smem_bytes, st>>>(tq, tk, tv, p);
  ctx->launches++;
  {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      cudaFuncAttributes fa;
      cudaFuncGetAttributes(&fa, attention_fwd_tc_kernel);
      set_error("attention_fwd_tc launch failed: %s (regs %d, max threads/block %d, static smem %zu, dynamic %d, local %zu)",
                cudaGetErrorString(e), fa.numRegs, fa.maxThreadsPerBlock, fa.sharedSizeBytes, smem_bytes, fa.localSizeBytes);
      return SIMSEG_ERR_CUDA;
    }
  }
  return SIMSEG_OK;
}

}  // namespace simseg
