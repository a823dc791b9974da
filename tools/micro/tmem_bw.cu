// TMEM read bandwidth probe: NW warps issue tcgen05.ld 32x32b.x32 back to back; prints bytes/clk/SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int UNROLL>
__global__ void probe(int iters, long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      uint32_t r[32];
      const uint32_t addr = base + ((i * UNROLL + u) & 15) * 32;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
          "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
            "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
            "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
            "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(addr) : "memory");
      if (u == UNROLL - 1) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc ^= r[0] ^ r[31];
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}
int main() {
  long long* d; uint32_t* s;
  cudaMalloc(&d, 148 * 8); cudaMalloc(&s, 4);
  const int iters = 2000;
  for (int nw : {1, 2, 4, 8, 16}) {
    probe<4><<<148, nw * 32>>>(iters, d, s);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const double bytes = (double)nw * 32 * 32 * 4 * 4 * iters;   // per CTA
    printf("warps %2d: %s  %lld clks  %.1f B/clk/SM (x32 loads, 4 in flight per warp)\n", nw, cudaGetErrorString(e), h[0], bytes / (double)h[0]);
  }
  return 0;
}
