// FP32 pipe throughput probe: FFMA (3-register), packed FFMA2 (fma.rn.f32x2), FMUL, MUFU.EX2, MUFU.RCP issued back to back from
// NW warps of one CTA with 8 independent chains per thread; prints lane-operations per clock per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_pipes fp32_pipes.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum { FFMA = 0, FFMA2 = 1, FMUL = 2, EX2 = 3, RCP = 4, FMUL2 = 5, MIX = 6 };

template <int OP>
__global__ void probe(int iters, long long* out, float* sink, float a, float b) {
  float x[8];
  unsigned long long p[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    x[i] = threadIdx.x * 1e-3f + i;
    p[i] = (static_cast<unsigned long long>(__float_as_uint(x[i])) << 32) | __float_as_uint(x[i] + 0.5f);
  }
  const unsigned long long pa = (static_cast<unsigned long long>(__float_as_uint(a)) << 32) | __float_as_uint(a);
  const unsigned long long pb = (static_cast<unsigned long long>(__float_as_uint(b)) << 32) | __float_as_uint(b);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
        if (OP == FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
        if (OP == FMUL) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(a));
        if (OP == FMUL2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pa));
        if (OP == EX2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        if (OP == RCP) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        if (OP == MIX) {      // 1 MUFU : 6 FFMA2, the shape of a polynomial epilogue
          if (i == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[0]));
          else asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
        }
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i] + __uint_as_float(static_cast<uint32_t>(p[i])) + __uint_as_float(static_cast<uint32_t>(p[i] >> 32));
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int lanes_per_instr) {
  long long* d_out;
  float* d_sink;
  cudaMalloc(&d_out, 148 * sizeof(long long));
  cudaMalloc(&d_sink, 148 * 1024 * sizeof(float));
  const int iters = 2048;
  for (int nw : {4, 8, 16, 32}) {
    probe<OP><<<148, nw * 32, 0>>>(iters, d_out, d_sink, 1.0001f, 1e-6f);
    cudaDeviceSynchronize();
    probe<OP><<<148, nw * 32, 0>>>(iters, d_out, d_sink, 1.0001f, 1e-6f);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    const double instr = static_cast<double>(iters) * 32 * nw;      // warp-instructions per SM
    printf("%-6s warps=%2d  %7.2f warp-instr/clk/SM  %7.1f lane-ops/clk/SM  (%lld clk)\n", name, nw, instr / h[0],
           instr * lanes_per_instr / h[0], h[0]);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
  cudaFree(d_out);
  cudaFree(d_sink);
}

int main() {
  run<FFMA>("FFMA", 32);
  run<FFMA2>("FFMA2", 64);
  run<FMUL>("FMUL", 32);
  run<FMUL2>("FMUL2", 64);
  run<EX2>("EX2", 32);
  run<RCP>("RCP", 32);
  run<MIX>("MIX1:7", 32);
  return 0;
}
