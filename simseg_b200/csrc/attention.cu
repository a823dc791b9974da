// Fused softmax(Q K^T * scale [+ key mask]) V, head_dim 64, short sequences (ViT: 197 / 325 tokens,
// BERT: 25 / 77 tokens): forward with log-sum-exp, and a two-phase backward that never materialises
// the S x S matrix.  Round-1 implementation on the warp-level tensor-core path (mma.sync m16n8k16,
// ldmatrix, cp.async); whole K/V (and Q/dO in backward) of one (batch, head) live in shared memory.
#include "common.cuh"

namespace simseg {

constexpr int HD = 64;                 // head dim
constexpr int ROWB = HD * 2;           // bytes per smem row

__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;       // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// smem tile [rows][64] bf16, 16-byte chunks XOR-swizzled by (row & 7): conflict-free ldmatrix
__device__ __forceinline__ uint32_t sw_addr(uint32_t base, int row, int chunk) {
  return base + row * ROWB + ((chunk ^ (row & 7)) << 4);
}

// cooperative load of `rows` rows (64 bf16 each) of one head into a swizzled smem tile; rows >= valid zero-filled
__device__ __forceinline__ void load_head_tile(uint32_t sbase, const __nv_bfloat16* g, int64_t stride_s, int rows, int valid) {
  for (int i = threadIdx.x; i < rows * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    const bool ok = r < valid;
    cp_async16(sw_addr(sbase, r, c), g + static_cast<int64_t>(ok ? r : 0) * stride_s + c * 8, ok);
  }
}

// A fragments (16 rows x 16 k) of a [row][64] tile: rows r0..r0+15, k-step ks
__device__ __forceinline__ void lda_frag(uint32_t (&a)[4], uint32_t sbase, int r0, int ks, int lane) {
  const int row = r0 + (lane & 7) + ((lane >> 3) & 1) * 8;
  const int chunk = ks * 2 + (lane >> 4);
  ldsm_x4(a, sw_addr(sbase, row, chunk));
}
// B fragments for two n-tiles from a tile stored [n][k] (non-transposed): n rows n0..n0+15, k-step ks
//   r[0],r[1] = (b0,b1) of n-tile n0 ; r[2],r[3] = (b0,b1) of n-tile n0+8
__device__ __forceinline__ void ldb_frag_nk(uint32_t (&r)[4], uint32_t sbase, int n0, int ks, int lane) {
  const int row = n0 + (lane & 7) + (lane >> 4) * 8;
  const int chunk = ks * 2 + ((lane >> 3) & 1);
  ldsm_x4(r, sw_addr(sbase, row, chunk));
}
// B fragments for two n-tiles from a tile stored [k][n] (transposed load): k rows k0..k0+15, n cols n0..n0+15
__device__ __forceinline__ void ldb_frag_kn(uint32_t (&r)[4], uint32_t sbase, int k0, int n0, int lane) {
  const int row = k0 + (lane & 7) + ((lane >> 3) & 1) * 8;
  const int chunk = (n0 >> 3) + (lane >> 4);
  ldsm_x4_t(r, sw_addr(sbase, row, chunk));
}

// ------------------------------------------------------------------------------------------------
// forward: CTA = (q-block of 16*warps rows, head, batch)
__global__ void __launch_bounds__(256) attention_fwd_kernel(
    const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v,
    int64_t stride_b, int64_t stride_s, int64_t stride_h, int H, int S, const int32_t* __restrict__ key_len,
    float scale_log2e, __nv_bfloat16* __restrict__ out, float* __restrict__ lse) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int b = blockIdx.z, h = blockIdx.y;
  const int warps = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int klen = key_len ? min(max(key_len[b], 1), S) : S;
  const int kpad = ((klen + 15) / 16) * 16;
  const uint32_t sK = smem_addr_u32(smem);
  const uint32_t sV = sK + ((S + 15) / 16) * 16 * ROWB;
  const int64_t base = static_cast<int64_t>(b) * stride_b + static_cast<int64_t>(h) * stride_h;
  load_head_tile(sK, k + base, stride_s, kpad, klen);
  load_head_tile(sV, v + base, stride_s, kpad, klen);

  const int q0 = (blockIdx.x * warps + warp) * 16;       // first query row of this warp
  // Q fragments straight from global memory (each row read once)
  uint32_t qa[4][4];
  {
    const int r_lo = q0 + (lane >> 2), r_hi = r_lo + 8;
    const __nv_bfloat16* qb = q + base;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int c = ks * 16 + (lane & 3) * 2;
      qa[ks][0] = r_lo < S ? *reinterpret_cast<const uint32_t*>(qb + static_cast<int64_t>(r_lo) * stride_s + c) : 0u;
      qa[ks][1] = r_hi < S ? *reinterpret_cast<const uint32_t*>(qb + static_cast<int64_t>(r_hi) * stride_s + c) : 0u;
      qa[ks][2] = r_lo < S ? *reinterpret_cast<const uint32_t*>(qb + static_cast<int64_t>(r_lo) * stride_s + c + 8) : 0u;
      qa[ks][3] = r_hi < S ? *reinterpret_cast<const uint32_t*>(qb + static_cast<int64_t>(r_hi) * stride_s + c + 8) : 0u;
    }
  }
  cp_async_wait_all();
  __syncthreads();
  if (q0 >= S) return;

  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;

  for (int kc = 0; kc < kpad; kc += 64) {
    const int npairs = min(4, (kpad - kc) / 16);           // 16-key groups in this chunk (warp-uniform)
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        if (np < npairs) {
          uint32_t bf[4];
          ldb_frag_nk(bf, sK, kc + np * 16, ks, lane);
          mma_bf16(s[2 * np], qa[ks], bf[0], bf[1]);
          mma_bf16(s[2 * np + 1], qa[ks], bf[2], bf[3]);
        }
      }
    }
    // scale, mask, running max
    float mx_lo = m_lo, mx_hi = m_hi;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int col = kc + nt * 8 + (lane & 3) * 2;
      const bool live = nt < 2 * npairs;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool ok = live && (col + (e & 1)) < klen;
        s[nt][e] = ok ? s[nt][e] * scale_log2e : -INFINITY;
      }
      mx_lo = fmaxf(mx_lo, fmaxf(s[nt][0], s[nt][1]));
      mx_hi = fmaxf(mx_hi, fmaxf(s[nt][2], s[nt][3]));
    }
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1)); mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1)); mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
    const float corr_lo = exp2f(m_lo - mx_lo), corr_hi = exp2f(m_hi - mx_hi);
    m_lo = mx_lo; m_hi = mx_hi;
    float rs_lo = 0.f, rs_hi = 0.f;
    uint32_t pa[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = exp2f(s[nt][0] - m_lo), p1 = exp2f(s[nt][1] - m_lo);
      const float p2 = exp2f(s[nt][2] - m_hi), p3 = exp2f(s[nt][3] - m_hi);
      rs_lo += p0 + p1; rs_hi += p2 + p3;
      pa[nt >> 1][(nt & 1) * 2] = pack_bf16(p0, p1);
      pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p2, p3);
    }
    l_lo = l_lo * corr_lo + rs_lo; l_hi = l_hi * corr_hi + rs_hi;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) { o[dt][0] *= corr_lo; o[dt][1] *= corr_lo; o[dt][2] *= corr_hi; o[dt][3] *= corr_hi; }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j < npairs) {
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {
          uint32_t bf[4];
          ldb_frag_kn(bf, sV, kc + j * 16, dp * 16, lane);
          mma_bf16(o[2 * dp], pa[j], bf[0], bf[1]);
          mma_bf16(o[2 * dp + 1], pa[j], bf[2], bf[3]);
        }
      }
    }
  }
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1); l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1); l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
  const float inv_lo = 1.0f / l_lo, inv_hi = 1.0f / l_hi;
  const int r_lo = q0 + (lane >> 2), r_hi = r_lo + 8;
  __nv_bfloat16* ob = out + (static_cast<int64_t>(b) * S) * (H * HD) + h * HD;
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) {
    const int c = dt * 8 + (lane & 3) * 2;
    if (r_lo < S) *reinterpret_cast<uint32_t*>(ob + static_cast<int64_t>(r_lo) * (H * HD) + c) = pack_bf16(o[dt][0] * inv_lo, o[dt][1] * inv_lo);
    if (r_hi < S) *reinterpret_cast<uint32_t*>(ob + static_cast<int64_t>(r_hi) * (H * HD) + c) = pack_bf16(o[dt][2] * inv_hi, o[dt][3] * inv_hi);
  }
  if (lse && (lane & 3) == 0) {
    float* lp = lse + (static_cast<int64_t>(b) * H + h) * S;
    if (r_lo < S) lp[r_lo] = (m_lo + log2f(l_lo)) * 0.69314718055994531f;
    if (r_hi < S) lp[r_hi] = (m_hi + log2f(l_hi)) * 0.69314718055994531f;
  }
}

// ------------------------------------------------------------------------------------------------
// backward: CTA = (head, batch).  smem: Q, K, V, dO tiles [Spad][64] + lse2[Spad] + D[Spad].
//   phase A: each warp owns 16-key tiles -> dK, dV (loops over all queries)
//   phase B: each warp owns 16-query tiles -> dQ  (loops over all keys)
__global__ void __launch_bounds__(256) attention_bwd_kernel(
    const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v,
    const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse,
    int64_t stride_b, int64_t stride_s, int64_t stride_h, int H, int S, const int32_t* __restrict__ key_len, float scale,
    __nv_bfloat16* __restrict__ dq, __nv_bfloat16* __restrict__ dk, __nv_bfloat16* __restrict__ dv) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int b = blockIdx.y, h = blockIdx.x;
  const int warps = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int klen = key_len ? min(max(key_len[b], 1), S) : S;
  const int spad = ((S + 15) / 16) * 16;
  const int kpad = ((klen + 15) / 16) * 16;
  const uint32_t sQ = smem_addr_u32(smem);
  const uint32_t sK = sQ + spad * ROWB;
  const uint32_t sV = sK + spad * ROWB;
  const uint32_t sdO = sV + spad * ROWB;
  float* s_lse2 = reinterpret_cast<float*>(smem + 4 * spad * ROWB);
  float* s_D = s_lse2 + spad;
  const int64_t base = static_cast<int64_t>(b) * stride_b + static_cast<int64_t>(h) * stride_h;
  const int64_t obase = (static_cast<int64_t>(b) * S) * (H * HD) + h * HD;
  const int64_t ostride = H * HD;
  load_head_tile(sQ, q + base, stride_s, spad, S);
  load_head_tile(sK, k + base, stride_s, kpad, klen);
  load_head_tile(sV, v + base, stride_s, kpad, klen);
  load_head_tile(sdO, dout + obase, ostride, spad, S);
  // D_i = sum_d dO[i,d] * O[i,d]; lse in log2 units
  for (int i = warp; i < spad; i += warps) {
    float d = 0.f;
    if (i < S) {
      const uint32_t a = *reinterpret_cast<const uint32_t*>(dout + obase + static_cast<int64_t>(i) * ostride + lane * 2);
      const uint32_t c = *reinterpret_cast<const uint32_t*>(out + obase + static_cast<int64_t>(i) * ostride + lane * 2);
      d = bf16_lo(a) * bf16_lo(c) + bf16_hi(a) * bf16_hi(c);
    }
    d = warp_sum(d);
    if (lane == 0) {
      s_D[i] = d;
      s_lse2[i] = i < S ? lse[(static_cast<int64_t>(b) * H + h) * S + i] * 1.44269504088896341f : 0.f;
    }
  }
  cp_async_wait_all();
  __syncthreads();
  const float scale_log2e = scale * 1.44269504088896341f;

  // ------------------------------ phase A: dK, dV ------------------------------
  for (int kt = warp * 16; kt < kpad; kt += warps * 16) {
    uint32_t ka[4][4], va[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) { lda_frag(ka[ks], sK, kt, ks, lane); lda_frag(va[ks], sV, kt, ks, lane); }
    float dkacc[8][4], dvacc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { dkacc[i][0] = dkacc[i][1] = dkacc[i][2] = dkacc[i][3] = 0.f; dvacc[i][0] = dvacc[i][1] = dvacc[i][2] = dvacc[i][3] = 0.f; }
    const int key_lo = kt + (lane >> 2), key_hi = key_lo + 8;
    for (int qc = 0; qc < spad; qc += 32) {
      const int npairs = min(2, (spad - qc) / 16);
      float st[4][4], dpt[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f; dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          if (np < npairs) {
            uint32_t bf[4];
            ldb_frag_nk(bf, sQ, qc + np * 16, ks, lane);          // S^T = K Q^T
            mma_bf16(st[2 * np], ka[ks], bf[0], bf[1]);
            mma_bf16(st[2 * np + 1], ka[ks], bf[2], bf[3]);
            ldb_frag_nk(bf, sdO, qc + np * 16, ks, lane);         // dP^T = V dO^T
            mma_bf16(dpt[2 * np], va[ks], bf[0], bf[1]);
            mma_bf16(dpt[2 * np + 1], va[ks], bf[2], bf[3]);
          }
        }
      }
      uint32_t pa[2][4], dsa[2][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int qi = qc + nt * 8 + (lane & 3) * 2;
        float p[4], ds[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int qq = qi + (e & 1);
          const int kk = (e < 2) ? key_lo : key_hi;
          const bool ok = (nt < 2 * npairs) && kk < klen && qq < S;
          const float pv = ok ? exp2f(st[nt][e] * scale_log2e - s_lse2[min(qq, spad - 1)]) : 0.f;
          p[e] = pv;
          ds[e] = pv * (dpt[nt][e] - s_D[min(qq, spad - 1)]);
        }
        pa[nt >> 1][(nt & 1) * 2] = pack_bf16(p[0], p[1]);
        pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p[2], p[3]);
        dsa[nt >> 1][(nt & 1) * 2] = pack_bf16(ds[0], ds[1]);
        dsa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(ds[2], ds[3]);
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (j < npairs) {
#pragma unroll
          for (int dp = 0; dp < 4; ++dp) {
            uint32_t bf[4];
            ldb_frag_kn(bf, sdO, qc + j * 16, dp * 16, lane);     // dV += P^T dO
            mma_bf16(dvacc[2 * dp], pa[j], bf[0], bf[1]);
            mma_bf16(dvacc[2 * dp + 1], pa[j], bf[2], bf[3]);
            ldb_frag_kn(bf, sQ, qc + j * 16, dp * 16, lane);      // dK += dS^T Q
            mma_bf16(dkacc[2 * dp], dsa[j], bf[0], bf[1]);
            mma_bf16(dkacc[2 * dp + 1], dsa[j], bf[2], bf[3]);
          }
        }
      }
    }
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      const int c = dt * 8 + (lane & 3) * 2;
      if (key_lo < S) {
        *reinterpret_cast<uint32_t*>(dk + base + static_cast<int64_t>(key_lo) * stride_s + c) = pack_bf16(dkacc[dt][0] * scale, dkacc[dt][1] * scale);
        *reinterpret_cast<uint32_t*>(dv + base + static_cast<int64_t>(key_lo) * stride_s + c) = pack_bf16(dvacc[dt][0], dvacc[dt][1]);
      }
      if (key_hi < S) {
        *reinterpret_cast<uint32_t*>(dk + base + static_cast<int64_t>(key_hi) * stride_s + c) = pack_bf16(dkacc[dt][2] * scale, dkacc[dt][3] * scale);
        *reinterpret_cast<uint32_t*>(dv + base + static_cast<int64_t>(key_hi) * stride_s + c) = pack_bf16(dvacc[dt][2], dvacc[dt][3]);
      }
    }
  }
  // keys in [kpad, S) (fully masked tiles) get zero gradients
  for (int i = threadIdx.x; i < (S - min(kpad, S)) * 8; i += blockDim.x) {
    const int r = kpad + (i >> 3), c = (i & 7) * 8;
    *reinterpret_cast<uint4*>(dk + base + static_cast<int64_t>(r) * stride_s + c) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(dv + base + static_cast<int64_t>(r) * stride_s + c) = make_uint4(0, 0, 0, 0);
  }

  // ------------------------------ phase B: dQ ------------------------------
  for (int qt = warp * 16; qt < spad; qt += warps * 16) {
    uint32_t qa[4][4], doa[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) { lda_frag(qa[ks], sQ, qt, ks, lane); lda_frag(doa[ks], sdO, qt, ks, lane); }
    float dqacc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { dqacc[i][0] = dqacc[i][1] = dqacc[i][2] = dqacc[i][3] = 0.f; }
    const int q_lo = qt + (lane >> 2), q_hi = q_lo + 8;
    const float l2_lo = s_lse2[q_lo], l2_hi = s_lse2[q_hi];
    const float D_lo = s_D[q_lo], D_hi = s_D[q_hi];
    for (int kc = 0; kc < kpad; kc += 32) {
      const int npairs = min(2, (kpad - kc) / 16);
      float s[4][4], dp_[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; dp_[i][0] = dp_[i][1] = dp_[i][2] = dp_[i][3] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          if (np < npairs) {
            uint32_t bf[4];
            ldb_frag_nk(bf, sK, kc + np * 16, ks, lane);          // S = Q K^T
            mma_bf16(s[2 * np], qa[ks], bf[0], bf[1]);
            mma_bf16(s[2 * np + 1], qa[ks], bf[2], bf[3]);
            ldb_frag_nk(bf, sV, kc + np * 16, ks, lane);          // dP = dO V^T
            mma_bf16(dp_[2 * np], doa[ks], bf[0], bf[1]);
            mma_bf16(dp_[2 * np + 1], doa[ks], bf[2], bf[3]);
          }
        }
      }
      uint32_t dsa[2][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int kj = kc + nt * 8 + (lane & 3) * 2;
        float ds[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const bool ok = (nt < 2 * npairs) && (kj + (e & 1)) < klen;
          const float pv = ok ? exp2f(s[nt][e] * scale_log2e - (e < 2 ? l2_lo : l2_hi)) : 0.f;
          ds[e] = pv * (dp_[nt][e] - (e < 2 ? D_lo : D_hi));
        }
        dsa[nt >> 1][(nt & 1) * 2] = pack_bf16(ds[0], ds[1]);
        dsa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(ds[2], ds[3]);
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (j < npairs) {
#pragma unroll
          for (int dp = 0; dp < 4; ++dp) {
            uint32_t bf[4];
            ldb_frag_kn(bf, sK, kc + j * 16, dp * 16, lane);      // dQ += dS K
            mma_bf16(dqacc[2 * dp], dsa[j], bf[0], bf[1]);
            mma_bf16(dqacc[2 * dp + 1], dsa[j], bf[2], bf[3]);
          }
        }
      }
    }
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      const int c = dt * 8 + (lane & 3) * 2;
      if (q_lo < S) *reinterpret_cast<uint32_t*>(dq + base + static_cast<int64_t>(q_lo) * stride_s + c) = pack_bf16(dqacc[dt][0] * scale, dqacc[dt][1] * scale);
      if (q_hi < S) *reinterpret_cast<uint32_t*>(dq + base + static_cast<int64_t>(q_hi) * stride_s + c) = pack_bf16(dqacc[dt][2] * scale, dqacc[dt][3] * scale);
    }
  }
}

// ------------------------------------------------------------------------------------------------
static int pick_warps(int tiles) {
  if (tiles <= 8) return tiles;
  const int ctas = static_cast<int>(cdiv(tiles, 8));
  return static_cast<int>(cdiv(tiles, ctas));
}

int attention_fwd_impl(Ctx* ctx, const void* q, const void* k, const void* v, int64_t stride_b, int64_t stride_s,
                       int64_t stride_h, int B, int H, int S, const int32_t* key_len, float scale, void* out, float* lse,
                       cudaStream_t st) {
  SIMSEG_CHECK_ARG(B > 0 && H > 0 && S > 0 && S <= 640, "attention_fwd: S=%d unsupported (1..640)", S);
  SIMSEG_CHECK_ARG(stride_s % 8 == 0 && stride_h % 8 == 0 && stride_b % 8 == 0, "attention: strides must be multiples of 8 elements");
  const int tiles = (S + 15) / 16;
  const int warps = pick_warps(tiles);
  const int smem_bytes = 2 * tiles * 16 * ROWB;
  static int max_set = 0;
  if (smem_bytes > max_set) {
    SIMSEG_CUDA(cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    max_set = smem_bytes;
  }
  dim3 grid(static_cast<unsigned>(cdiv(tiles, warps)), H, B);
  attention_fwd_kernel<<<grid, warps * 32, smem_bytes, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(q), reinterpret_cast<const __nv_bfloat16*>(k),
      reinterpret_cast<const __nv_bfloat16*>(v), stride_b, stride_s, stride_h, H, S, key_len,
      scale * 1.44269504088896341f, reinterpret_cast<__nv_bfloat16*>(out), lse);
  ctx->launches++;
  SIMSEG_LAUNCH_CHECK();
  return SIMSEG_OK;
}

int attention_bwd_impl(Ctx* ctx, const void* q, const void* k, const void* v, const void* out, const void* dout,
                       const float* lse, int64_t stride_b, int64_t stride_s, int64_t stride_h, int B, int H, int S,
                       const int32_t* key_len, float scale, void* dq, void* dk, void* dv, cudaStream_t st) {
  SIMSEG_CHECK_ARG(B > 0 && H > 0 && S > 0 && S <= 400, "attention_bwd: S=%d unsupported (1..400)", S);
  SIMSEG_CHECK_ARG(stride_s % 8 == 0 && stride_h % 8 == 0 && stride_b % 8 == 0, "attention: strides must be multiples of 8 elements");
  const int tiles = (S + 15) / 16;
  const int warps = pick_warps(tiles);
  const int spad = tiles * 16;
  const int smem_bytes = 4 * spad * ROWB + 2 * spad * static_cast<int>(sizeof(float));
  static int max_set = 0;
  if (smem_bytes > max_set) {
    SIMSEG_CUDA(cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    max_set = smem_bytes;
  }
  dim3 grid(H, B);
  attention_bwd_kernel<<<grid, warps * 32, smem_bytes, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(q), reinterpret_cast<const __nv_bfloat16*>(k),
      reinterpret_cast<const __nv_bfloat16*>(v), reinterpret_cast<const __nv_bfloat16*>(out),
      reinterpret_cast<const __nv_bfloat16*>(dout), lse, stride_b, stride_s, stride_h, H, S, key_len, scale,
      reinterpret_cast<__nv_bfloat16*>(dq), reinterpret_cast<__nv_bfloat16*>(dk), reinterpret_cast<__nv_bfloat16*>(dv));
  ctx->launches++;
  SIMSEG_LAUNCH_CHECK();
  return SIMSEG_OK;
}

}  // namespace simseg
