// Bandwidth-bound row / elementwise kernels: casts, column sums, GELU recompute, LayerNorm fwd/bwd.
// All use 128-bit global accesses; LayerNorm keeps a whole row in the registers of one warp.
#include "common.cuh"
#include "philox.cuh"

namespace simseg {

// ------------------------------------------------------------------------------------------------
// fp32 -> bf16 cast (+ optional transposed copy) for weights.  32x32 smem tile, coalesced both ways.
__global__ void cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                 __nv_bfloat16* __restrict__ dst_t, int64_t rows, int64_t cols) {
  __shared__ float tile[32][33];
  const int64_t c = blockIdx.x * 32 + threadIdx.x;
  const int64_t r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int64_t r = r0 + i;
    float v = 0.f;
    if (r < rows && c < cols) {
      v = src[r * cols + c];
      if (dst) dst[r * cols + c] = __float2bfloat16(v);
    }
    tile[i][threadIdx.x] = v;
  }
  if (!dst_t) return;
  __syncthreads();
  const int64_t tr = blockIdx.x * 32;     // transposed: row index = original column
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int64_t oc = tr + i;            // original column
    const int64_t orow = r0 + threadIdx.x;
    if (oc < cols && orow < rows) dst_t[oc * rows + orow] = __float2bfloat16(tile[threadIdx.x][i]);
  }
}

int cast_bf16_impl(Ctx* ctx, const float* src, void* dst, void* dst_t, int64_t rows, int64_t cols, cudaStream_t st) {
  SIMSEG_CHECK_ARG(rows > 0 && cols > 0, "cast_bf16: empty");
  dim3 grid(static_cast<unsigned>(cdiv(cols, 32)), static_cast<unsigned>(cdiv(rows, 32)));
  cast_bf16_kernel<<<grid, dim3(32, 8), 0, st>>>(src, reinterpret_cast<__nv_bfloat16*>(dst),
                                                 reinterpret_cast<__nv_bfloat16*>(dst_t), rows, cols);
  ctx->launches++;
  SIMSEG_LAUNCH_CHECK();
  return SIMSEG_OK;
}

constexpr int kCastTableMax = 512;
constexpr int kCastSub = 4;                           // 32 x 32 sub-tiles per block (32 rows x 128 columns)
// Many weights in ONE launch (the per-step bf16 refresh of every Linear weight: ~250 tiny launches otherwise).  The item
// table lives in device memory; block -> item by binary search over the items' first block index (a shared-memory copy of
// that column).  A block converts 32 rows x 128 columns: with one 32 x 32 tile per block the refresh was paced by the
// dependent chain at the head of every block (table -> item -> tile loads, ~3 us for 4 KB: 2.0 TB/s over 86 k blocks);
// four sub-tiles put 16 loads per thread in flight behind one such chain.
__global__ void __launch_bounds__(256) cast_bf16_multi_kernel(const simseg_cast_item* __restrict__ items, int n_items) {
  __shared__ float tile[kCastSub][32][33];
  __shared__ int64_t first[kCastTableMax];
  const int64_t blk = blockIdx.x;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const bool in_smem = n_items <= kCastTableMax;
  if (in_smem) {
    for (int i = tid; i < n_items; i += 256) first[i] = items[i].first_block;
    __syncthreads();
  }
  int lo = 0, hi = n_items - 1;
  while (lo < hi) {                                   // last item with first_block <= blk
    const int mid = (lo + hi + 1) >> 1;
    if ((in_smem ? first[mid] : items[mid].first_block) <= blk) lo = mid; else hi = mid - 1;
  }
  const simseg_cast_item it = items[lo];
  const int64_t lb = blk - it.first_block;
  const int64_t bx_n = (it.cols + 32 * kCastSub - 1) / (32 * kCastSub);
  const int64_t bx = lb % bx_n, by = lb / bx_n;
  const float* src = it.src;
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(it.dst);
  __nv_bfloat16* dst_t = reinterpret_cast<__nv_bfloat16*>(it.dst_t);
  const int64_t r0 = by * 32;
  float v[kCastSub][4];
#pragma unroll
  for (int s = 0; s < kCastSub; ++s) {
    const int64_t c = (bx * kCastSub + s) * 32 + threadIdx.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t r = r0 + threadIdx.y + 8 * k;
      v[s][k] = (r < it.rows && c < it.cols) ? __ldg(src + r * it.cols + c) : 0.f;
    }
  }
#pragma unroll
  for (int s = 0; s < kCastSub; ++s) {
    const int64_t c = (bx * kCastSub + s) * 32 + threadIdx.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = threadIdx.y + 8 * k;
      const int64_t r = r0 + i;
      if (dst && r < it.rows && c < it.cols) dst[r * it.ld + c] = __float2bfloat16(v[s][k]);
      tile[s][i][threadIdx.x] = v[s][k];
    }
  }
  if (!dst_t) return;
  __syncthreads();
#pragma unroll
  for (int s = 0; s < kCastSub; ++s) {
    for (int i = threadIdx.y; i < 32; i += 8) {
      const int64_t oc = (bx * kCastSub + s) * 32 + i;  // original column = transposed row
      const int64_t orow = r0 + threadIdx.x;
      if (oc < it.cols && orow < it.rows) dst_t[oc * it.ld_t + orow] = __float2bfloat16(tile[s][threadIdx.x][i]);
    }
  }
}

int cast_bf16_multi_impl(Ctx* ctx, const simseg_cast_item* items_dev, int n_items, int64_t total_blocks, cudaStream_t st) {
  SIMSEG_CHECK_ARG(items_dev != nullptr && n_items > 0 && total_blocks > 0 && total_blocks < (int64_t(1) << 31),
                   "cast_bf16_multi: bad table (n_items=%d, blocks=%lld)", n_items, static_cast<long long>(total_blocks));
  cast_bf16_multi_kernel<<<static_cast<unsigned>(total_blocks), dim3(32, 8), 0, st>>>(items_dev, n_items);
  ctx->launches++;
  SIMSEG_LAUNCH_CHECK();
  return SIMSEG_OK;
}

// ------------------------------------------------------------------------------------------------
// column sums: out[n] += sum_m x[m,n].  Block = 32 column-lanes x 8 row-lanes, each lane owns 4 columns.
// bf16: 8 columns (one 16-byte load) per thread, 4 rows in flight per thread; fp32: 4 columns per thread.
template <bool BF16>
__global__ void __launch_bounds__(256) colsum_kernel(const void* __restrict__ x, int64_t M, int64_t N, int64_t ldx,
                                                     float* __restrict__ out, int64_t rows_per_block) {
  constexpr int CPT = BF16 ? 8 : 4;                       // columns per thread
  const int64_t c0 = (static_cast<int64_t>(blockIdx.x) * 32 + threadIdx.x) * CPT;
  const int64_t r_begin = blockIdx.y * rows_per_block;
  const int64_t r_end = min(M, r_begin + rows_per_block);
  float a[CPT];
#pragma unroll
  for (int j = 0; j < CPT; ++j) a[j] = 0.f;
  if (c0 < N) {
    const bool full = c0 + CPT <= N;
    if (full) {
      int64_t r = r_begin + threadIdx.y;
      for (; r + 24 < r_end; r += 32) {                   // four independent 16-byte loads in flight
        uint4 u[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const char* pp = reinterpret_cast<const char*>(x) + ((r + 8 * k) * ldx + c0) * (BF16 ? 2 : 4);
          u[k] = ldg_nc_v4(pp);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (BF16) {
            a[0] += bf16_lo(u[k].x); a[1] += bf16_hi(u[k].x); a[2] += bf16_lo(u[k].y); a[3] += bf16_hi(u[k].y);
            a[4 % CPT] += BF16 ? bf16_lo(u[k].z) : 0.f; a[5 % CPT] += BF16 ? bf16_hi(u[k].z) : 0.f;
            a[6 % CPT] += BF16 ? bf16_lo(u[k].w) : 0.f; a[7 % CPT] += BF16 ? bf16_hi(u[k].w) : 0.f;
          } else {
            a[0] += __uint_as_float(u[k].x); a[1] += __uint_as_float(u[k].y);
            a[2] += __uint_as_float(u[k].z); a[3] += __uint_as_float(u[k].w);
          }
        }
      }
      for (; r < r_end; r += 8) {
        const uint4 u = ldg_nc_v4(reinterpret_cast<const char*>(x) + (r * ldx + c0) * (BF16 ? 2 : 4));
        if (BF16) {
          a[0] += bf16_lo(u.x); a[1] += bf16_hi(u.x); a[2] += bf16_lo(u.y); a[3] += bf16_hi(u.y);
          a[4 % CPT] += BF16 ? bf16_lo(u.z) : 0.f; a[5 % CPT] += BF16 ? bf16_hi(u.z) : 0.f;
          a[6 % CPT] += BF16 ? bf16_lo(u.w) : 0.f; a[7 % CPT] += BF16 ? bf16_hi(u.w) : 0.f;
        } else {
          a[0] += __uint_as_float(u.x); a[1] += __uint_as_float(u.y); a[2] += __uint_as_float(u.z); a[3] += __uint_as_float(u.w);
        }
      }
    } else {
      for (int64_t r = r_begin + threadIdx.y; r < r_end; r += 8)
        for (int j = 0; j < CPT; ++j)
          if (c0 + j < N)
            a[j] += BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[r * ldx + c0 + j])
                         : reinterpret_cast<const float*>(x)[r * ldx + c0 + j];
    }
  }
  __shared__ float red[8][32][CPT + 1];
#pragma unroll
  for (int j = 0; j < CPT; ++j) red[threadIdx.y][threadIdx.x][j] = a[j];
  __syncthreads();
  if (threadIdx.y == 0 && c0 < N) {
    for (int j = 0; j < CPT; ++j) {
      float s = 0.f;
      for (int y = 0; y < 8; ++y) s += red[y][threadIdx.x][j];
      if (c0 + j < N) atomicAdd(out + c0 + j, s);
    }
  }
}

int colsum_impl(Ctx* ctx, const void* x, int dtype, int64_t M, int64_t N, int64_t ldx, float* out, int accumulate,
                cudaStream_t st) {
  SIMSEG_CHECK_ARG(M > 0 && N > 0, "colsum: empty");
  const int eb = dtype == SIMSEG_BF16 ? 2 : 4;
  SIMSEG_CHECK_ARG((ldx * eb) % 16 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "colsum: rows must be 16B aligned");
  if (!accumulate) SIMSEG_CUDA(cudaMemsetAsync(out, 0, N * sizeof(float), st));
  const int64_t col_blocks = cdiv(N, dtype == SIMSEG_BF16 ? 256 : 128);
  int64_t row_blocks = cdiv(static_cast<int64_t>(ctx->num_sms) * 8, col_blocks);
  const int64_t max_rb = cdiv(M, 64);
  if (row_blocks > max_rb) row_blocks = max_rb;
  if (row_blocks < 1) row_blocks = 1;
  const int64_t rpb = cdiv(M, row_blocks);
  dim3 grid(static_cast<unsigned>(col_blocks), static_cast<unsigned>(cdiv(M, rpb)));
  if (dtype == SIMSEG_BF16) colsum_kernel<true><<<grid, dim3(32, 8), 0, st>>>(x, M, N, ldx, out, rpb);
  else colsum_kernel<false><<<grid, dim3(32, 8), 0, st>>>(x, M, N, ldx, out, rpb);
  ctx->launches++;
  SIMSEG_LAUNCH_CHECK();
  return SIMSEG_OK;
}

// ------------------------------------------------------------------------------------------------
__global__ void gelu_fwd_kernel(const uint4* __restrict__ h, uint4* __restrict__ a, int64_t n8) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint4 u = ldg_nc_v4(h + i);
    uint4 o;
    o.x = pack_bf16(gelu_erf(bf16_lo(u.x)), gelu_erf(bf16_hi(u.x)));
    o.y = pack_bf16(gelu_erf(bf16_lo(u.y)), gelu_erf(bf16_hi(u.y)));
    o.z = pack_bf16(gelu_erf(bf16_lo(u.z)), gelu_erf(bf16_hi(u.z)));
    o.w = pack_bf16(gelu_erf(bf16_lo(u.w)), gelu_erf(bf16_hi(u.w)));
    a[i] = o;
  }
}

int gelu_fwd_impl(Ctx* ctx, const void* h, void* a, int64_t n, cudaStream_t st) {
  SIMSEG_CHECK_ARG(n > 0 && n % 8 == 0, "gelu_fwd: n must be a positive multiple of 8");
  const int64_t n8 = n / 8;
  const int grid = static_cast<int>(imin64(cdiv(n8, 256), static_cast<int64_t>(ctx->num_sms) * 16));
  gelu_fwd_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const uint4*>(h), reinterpret_cast<uint4*>(a), n8);
  ctx->launches++;
  SIMSEG_LAUNCH_CHECK();
  return SIMSEG_OK;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm.  One warp per row; lane owns V float4 groups: columns (g*32 + lane)*4 .. +3, g < V (D = 128 V).
template <int V, bool XBF16>
__device__ __forceinline__ void load_row(const void* x, int64_t row, int D, int lane, float (&v)[V][4]) {
#pragma unroll
  for (int g = 0; g < V; ++g) {
    const int c = (g * 32 + lane) * 4;
    if (XBF16) {
      const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(x) + row * D + c);
      v[g][0] = bf16_lo(u.x); v[g][1] = bf16_hi(u.x); v[g][2] = bf16_lo(u.y); v[g][3] = bf16_hi(u.y);
    } else {
      const float4 f = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + row * D + c);
      v[g][0] = f.x; v[g][1] = f.y; v[g][2] = f.z; v[g][3] = f.w;
    }
  }
}

// DROP (train-mode BERT dropout, philox.cuh): 0 none; 1 the `add` summand is dropped (BertSelfOutput / BertOutput:
// LayerNorm(dropout(dense(..)) + input)); 2 the outputs are dropped (BertEmbeddings: dropout(LayerNorm(e))).
template <int V, bool XBF16, int DROP>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const void* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps, int64_t M,
                                                            __nv_bfloat16* __restrict__ y_bf16, float* __restrict__ y_f32,
                                                            float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                            const __nv_bfloat16* __restrict__ add, float* __restrict__ sum_out,
                                                            const DropSpec ds) {
  constexpr int D = V * 128;
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  DropKey dk{};
  if (DROP) dk = load_drop_key(ds);
  float g[V][4], b[V][4];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float4 gg = *reinterpret_cast<const float4*>(gamma + (i * 32 + lane) * 4);
    const float4 bb = *reinterpret_cast<const float4*>(beta + (i * 32 + lane) * 4);
    g[i][0] = gg.x; g[i][1] = gg.y; g[i][2] = gg.z; g[i][3] = gg.w;
    b[i][0] = bb.x; b[i][1] = bb.y; b[i][2] = bb.z; b[i][3] = bb.w;
  }
  for (int64_t row = warp_global; row < M; row += nwarps) {
    float v[V][4];
    load_row<V, XBF16>(x, row, D, lane, v);
    if (add != nullptr) {
      // fused residual add: s = x + add (the bf16 output of the preceding Linear); s is the new residual stream
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const int c = (i * 32 + lane) * 4;
        const uint2 u = *reinterpret_cast<const uint2*>(add + row * D + c);
        if (DROP == 1) {
          const uint4 w = drop_words(ds, dk, static_cast<uint32_t>(i * 32 + lane), static_cast<uint32_t>(row));
          v[i][0] += w.x >= ds.thr ? bf16_lo(u.x) * ds.inv_keep : 0.f;
          v[i][1] += w.y >= ds.thr ? bf16_hi(u.x) * ds.inv_keep : 0.f;
          v[i][2] += w.z >= ds.thr ? bf16_lo(u.y) * ds.inv_keep : 0.f;
          v[i][3] += w.w >= ds.thr ? bf16_hi(u.y) * ds.inv_keep : 0.f;
        } else {
          v[i][0] += bf16_lo(u.x); v[i][1] += bf16_hi(u.x); v[i][2] += bf16_lo(u.y); v[i][3] += bf16_hi(u.y);
        }
        if (sum_out) *reinterpret_cast<float4*>(sum_out + row * D + c) = make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
      }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) s += (v[i][0] + v[i][1]) + (v[i][2] + v[i][3]);
    const float mu = warp_sum(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float d = v[i][j] - mu; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mu;
      if (rstd_out) rstd_out[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int c = (i * 32 + lane) * 4;
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = (v[i][j] - mu) * rstd * g[i][j] + b[i][j];
      if (DROP == 2) {
        const uint4 w = drop_words(ds, dk, static_cast<uint32_t>(i * 32 + lane), static_cast<uint32_t>(row));
        o[0] = w.x >= ds.thr ? o[0] * ds.inv_keep : 0.f;
        o[1] = w.y >= ds.thr ? o[1] * ds.inv_keep : 0.f;
        o[2] = w.z >= ds.thr ? o[2] * ds.inv_keep : 0.f;
        o[3] = w.w >= ds.thr ? o[3] * ds.inv_keep : 0.f;
      }
      if (y_bf16) {
        uint2 u; u.x = pack_bf16(o[0], o[1]); u.y = pack_bf16(o[2], o[3]);
        *reinterpret_cast<uint2*>(y_bf16 + row * D + c) = u;
      }
      if (y_f32) *reinterpret_cast<float4*>(y_f32 + row * D + c) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
}

// DROP as in the forward: 1 -> dx_bf16 and dx_colsum (gradient of the dropped summand and of its Linear's bias) carry the
// mask, dx (the residual path) does not; 2 -> the incoming gradient dy (+ dy2) is masked first.
template <int V, bool XBF16, bool DYBF16, int DROP>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(
    const void* __restrict__ dy, const float* __restrict__ dy2, const void* __restrict__ x, const float* __restrict__ gamma,
    const float* __restrict__ mean, const float* __restrict__ rstd, int64_t M, float* __restrict__ dx, int dx_accumulate,
    __nv_bfloat16* __restrict__ dx_bf16, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dx_colsum,
    const DropSpec ds) {
  constexpr int D = V * 128;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  DropKey dk{};
  if (DROP) dk = load_drop_key(ds);
  const int64_t warp_global = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  float g[V][4];
  float acc_g[V][4], acc_b[V][4], acc_x[V][4];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float4 gg = *reinterpret_cast<const float4*>(gamma + (i * 32 + lane) * 4);
    g[i][0] = gg.x; g[i][1] = gg.y; g[i][2] = gg.z; g[i][3] = gg.w;
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc_g[i][j] = 0.f; acc_b[i][j] = 0.f; acc_x[i][j] = 0.f; }
  }
  for (int64_t row = warp_global; row < M; row += nwarps) {
    float xv[V][4], dv[V][4];
    load_row<V, XBF16>(x, row, D, lane, xv);
    load_row<V, DYBF16>(dy, row, D, lane, dv);
    // the running residual gradient is requested together with x and dy (one HBM round trip per row, not two)
    float pv[V][4];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
      if (dx_accumulate) f = *reinterpret_cast<const float4*>(dx + row * D + (i * 32 + lane) * 4);
      pv[i][0] = f.x; pv[i][1] = f.y; pv[i][2] = f.z; pv[i][3] = f.w;
    }
    if (dy2) {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float4 f = *reinterpret_cast<const float4*>(dy2 + row * D + (i * 32 + lane) * 4);
        dv[i][0] += f.x; dv[i][1] += f.y; dv[i][2] += f.z; dv[i][3] += f.w;
      }
    }
    if (DROP == 2) {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const uint4 w = drop_words(ds, dk, static_cast<uint32_t>(i * 32 + lane), static_cast<uint32_t>(row));
        dv[i][0] = w.x >= ds.thr ? dv[i][0] * ds.inv_keep : 0.f;
        dv[i][1] = w.y >= ds.thr ? dv[i][1] * ds.inv_keep : 0.f;
        dv[i][2] = w.z >= ds.thr ? dv[i][2] * ds.inv_keep : 0.f;
        dv[i][3] = w.w >= ds.thr ? dv[i][3] * ds.inv_keep : 0.f;
      }
    }
    const float mu = mean[row], rs = rstd[row];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float xh = (xv[i][j] - mu) * rs;
        const float gy = dv[i][j] * g[i][j];
        c1 += gy; c2 += gy * xh;
        acc_g[i][j] += dv[i][j] * xh;
        acc_b[i][j] += dv[i][j];
        xv[i][j] = xh;
        dv[i][j] = gy;
      }
    c1 = warp_sum(c1) * (1.0f / D);
    c2 = warp_sum(c2) * (1.0f / D);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int c = (i * 32 + lane) * 4;
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = rs * (dv[i][j] - c1 - xv[i][j] * c2) + pv[i][j];
      if (dx) *reinterpret_cast<float4*>(dx + row * D + c) = make_float4(o[0], o[1], o[2], o[3]);
      if (DROP == 1) {
        const uint4 w = drop_words(ds, dk, static_cast<uint32_t>(i * 32 + lane), static_cast<uint32_t>(row));
        o[0] = w.x >= ds.thr ? o[0] * ds.inv_keep : 0.f;
        o[1] = w.y >= ds.thr ? o[1] * ds.inv_keep : 0.f;
        o[2] = w.z >= ds.thr ? o[2] * ds.inv_keep : 0.f;
        o[3] = w.w >= ds.thr ? o[3] * ds.inv_keep : 0.f;
      }
      if (dx_bf16) {
        uint2 u; u.x = pack_bf16(o[0], o[1]); u.y = pack_bf16(o[2], o[3]);
        *reinterpret_cast<uint2*>(dx_bf16 + row * D + c) = u;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) acc_x[i][j] += o[j];
    }
  }
  // block reduction of the per-column partials, then one atomic per column per block
  __shared__ float red[8][128 * V];
  auto reduce_store = [&](float (&a)[V][4], float* out) {
    if (out == nullptr) return;            // uniform across the block
    __syncthreads();
#pragma unroll
    for (int i = 0; i < V; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) red[warp][(i * 32 + lane) * 4 + j] = a[i][j];
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[w][c];
      atomicAdd(out + c, s);
    }
  };
  reduce_store(acc_g, dgamma);
  reduce_store(acc_b, dbeta);
  reduce_store(acc_x, dx_colsum);
}

static int ln_grid(Ctx* ctx, int64_t M) {
  const int64_t want = cdiv(M, 8);
  const int64_t cap = static_cast<int64_t>(ctx->num_sms) * 8;
  return static_cast<int>(want < cap ? want : cap);
}

int layernorm_fwd_impl(Ctx* ctx, const void* x, int x_dtype, const float* gamma, const float* beta, float eps, int64_t M,
                       int D, void* y_bf16, float* y_f32, float* mean, float* rstd, const void* add_v, float* sum_out,
                       float drop_p, const void* drop_rng, uint32_t drop_site, cudaStream_t st) {
  const __nv_bfloat16* add = reinterpret_cast<const __nv_bfloat16*>(add_v);
  SIMSEG_CHECK_ARG(M > 0, "layernorm_fwd: empty");
  SIMSEG_CHECK_ARG(D == 384 || D == 768 || D == 512 || D == 128 || D == 256, "layernorm: D=%d unsupported (128/256/384/512/768)", D);
  const int grid = ln_grid(ctx, M);
  auto* yb = reinterpret_cast<__nv_bfloat16*>(y_bf16);
  const bool drop = drop_p > 0.f;
  SIMSEG_CHECK_ARG(!drop || (drop_p < 1.f && drop_rng != nullptr && x_dtype == SIMSEG_F32 && M < (int64_t(1) << 32)),
                   "layernorm_fwd: dropout needs 0 < p < 1, a device {seed, step} pair and an fp32 x");
  const DropSpec ds = make_drop_spec(drop ? drop_p : 0.f, drop_rng, drop_site);
#define LN_FWD(V)                                                                                             \
  if (drop && add) layernorm_fwd_kernel<V, false, 1><<<grid, 256, 0, st>>>(x, gamma, beta, eps, M, yb, y_f32, mean, rstd, add, sum_out, ds); \
  else if (drop) layernorm_fwd_kernel<V, false, 2><<<grid, 256, 0, st>>>(x, gamma, beta, eps, M, yb, y_f32, mean, rstd, add, sum_out, ds); \
  else if (x_dtype == SIMSEG_BF16) layernorm_fwd_kernel<V, true, 0><<<grid, 256, 0, st>>>(x, gamma, beta, eps, M, yb, y_f32, mean, rstd, add, sum_out, ds); \
  else layernorm_fwd_kernel<V, false, 0><<<grid, 256, 0, st>>>(x, gamma, beta, eps, M, yb, y_f32, mean, rstd, add, sum_out, ds)
  switch (D / 128) {
    case 1: LN_FWD(1); break;
    case 2: LN_FWD(2); break;
    case 3: LN_FWD(3); break;
    case 4: LN_FWD(4); break;
    default: LN_FWD(6); break;
  }
#undef LN_FWD
  ctx->launches++;
  SIMSEG_LAUNCH_CHECK();
  return SIMSEG_OK;
}

int layernorm_bwd_impl(Ctx* ctx, const void* dy, int dy_dtype, const float* dy2, const void* x, int x_dtype,
                       const float* gamma, const float* mean, const float* rstd, int64_t M, int D, float* dx,
                       int dx_accumulate, void* dx_bf16, float* dgamma, float* dbeta, float* dx_colsum, int drop_mode,
                       float drop_p, const void* drop_rng, uint32_t drop_site, cudaStream_t st) {
  SIMSEG_CHECK_ARG(M > 0, "layernorm_bwd: empty");
  SIMSEG_CHECK_ARG(D == 384 || D == 768 || D == 512 || D == 128 || D == 256, "layernorm: D=%d unsupported", D);
  SIMSEG_CHECK_ARG(!(dx_accumulate && dx == nullptr), "layernorm_bwd: dx_accumulate needs dx");
  const int grid = ln_grid(ctx, M);
  auto* db = reinterpret_cast<__nv_bfloat16*>(dx_bf16);
  if (!(drop_p > 0.f)) drop_mode = 0;
  SIMSEG_CHECK_ARG(drop_mode == 0 || ((drop_mode == 1 || drop_mode == 2) && drop_p < 1.f && drop_rng != nullptr &&
                                      x_dtype == SIMSEG_F32 && M < (int64_t(1) << 32)),
                   "layernorm_bwd: dropout needs mode 1|2, 0 < p < 1, a device {seed, step} pair and an fp32 x");
  const DropSpec ds = make_drop_spec(drop_mode ? drop_p : 0.f, drop_rng, drop_site);
#define LN_BWD_K(V, XB, DB, DR) \
  layernorm_bwd_kernel<V, XB, DB, DR><<<grid, 256, 0, st>>>(dy, dy2, x, gamma, mean, rstd, M, dx, dx_accumulate, db, dgamma, dbeta, dx_colsum, ds)
#define LN_BWD(V)                                                                                                   \
  do {                                                                                                              \
    if (drop_mode == 1) {                                                                                           \
      if (dy_dtype == SIMSEG_BF16) LN_BWD_K(V, false, true, 1); else LN_BWD_K(V, false, false, 1);                 \
    } else if (drop_mode == 2) {                                                                                    \
      if (dy_dtype == SIMSEG_BF16) LN_BWD_K(V, false, true, 2); else LN_BWD_K(V, false, false, 2);                 \
    } else if (x_dtype == SIMSEG_BF16) {                                                                            \
      if (dy_dtype == SIMSEG_BF16) LN_BWD_K(V, true, true, 0); else LN_BWD_K(V, true, false, 0);                   \
    } else {                                                                                                        \
      if (dy_dtype == SIMSEG_BF16) LN_BWD_K(V, false, true, 0); else LN_BWD_K(V, false, false, 0);                 \
    }                                                                                                               \
  } while (0)
  switch (D / 128) {
    case 1: LN_BWD(1); break;
    case 2: LN_BWD(2); break;
    case 3: LN_BWD(3); break;
    case 4: LN_BWD(4); break;
    default: LN_BWD(6); break;
  }
#undef LN_BWD
#undef LN_BWD_K
  ctx->launches++;
  SIMSEG_LAUNCH_CHECK();
  return SIMSEG_OK;
}

}  // namespace simseg
