// Shared host/device helpers for libsimseg_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>

#include "../../include/simseg_b200.h"

namespace simseg {

// ---- error plumbing (thread-local message, negative return codes) -------------------------
void set_error(const char* fmt, ...);
#define SIMSEG_CHECK_ARG(cond, ...)                      \
  do {                                                   \
    if (!(cond)) {                                       \
      ::simseg::set_error(__VA_ARGS__);                  \
      return SIMSEG_ERR_INVALID;                         \
    }                                                    \
  } while (0)
#define SIMSEG_CUDA(call)                                                               \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      ::simseg::set_error("%s:%d CUDA error %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return SIMSEG_ERR_CUDA;                                                           \
    }                                                                                   \
  } while (0)
#define SIMSEG_LAUNCH_CHECK() SIMSEG_CUDA(cudaGetLastError())

struct Ctx {
  int device;
  int num_sms;
  int launches;            // kernels launched through this ctx (bench's gpu_launches claim)
};

inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }

// ---- device helpers ------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);   // .x = a (low half), .y = b
  return *reinterpret_cast<uint32_t*>(&t);
}

// erf-GELU (the reference's nn.GELU() default) and its derivative, fp32.  erf uses Abramowitz-Stegun 7.1.26
// (|abs err| <= 1.5e-7, far below the bf16 resolution of every consumer) so one MUFU.RCP + one MUFU.EX2 serve both
// gelu(x) = x*Phi(x) and gelu'(x) = Phi(x) + x*phi(x):  exp(-x^2/2) is shared between erf(x/sqrt2) and phi(x).
// Instruction counts matter here: these run inside GEMM epilogues that are ALU-issue-bound.
//   w(x) = 0.5 * erfc(|x|/sqrt2) = t*(a1' + t*(a2' + t*(a3' + t*(a4' + t*a5')))) * exp(-x^2/2),  t = 1/(1 + p|x|/sqrt2)
//   (coefficients pre-multiplied by 0.5);   Phi(x) = x >= 0 ? 1 - w : w
__device__ __forceinline__ float gelu_half_erfc(float x, float& e) {
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f * 0.70710678118654752f, fabsf(x), 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"((x * x) * (-0.5f * 1.44269504088896341f)));   // exp(-x^2/2)
  float q = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  q = fmaf(q, t, 0.5f * 1.421413741f);
  q = fmaf(q, t, 0.5f * -0.284496736f);
  q = fmaf(q, t, 0.5f * 0.254829592f);
  return (q * t) * e;
}
__device__ __forceinline__ void gelu_erf_both(float x, float& gelu, float& dgelu) {
  float e;
  const float w = gelu_half_erfc(x, e);
  const float cdf = x >= 0.f ? 1.0f - w : w;
  gelu = x * cdf;
  dgelu = fmaf(x * 0.3989422804014327f, e, cdf);
}
// gelu(x) = x*Phi(x) = max(x,0) - |x|*w(x): no select, no explicit Phi
__device__ __forceinline__ float gelu_erf(float x) {
  float e;
  const float w = gelu_half_erfc(x, e);
  return fmaf(-fabsf(x), w, fmaxf(x, 0.f));
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float g, d;
  gelu_erf_both(x, g, d);
  return d;
}

// 128-bit streaming loads/stores
__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_cs_v4(void* p, uint4 v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
#endif

}  // namespace simseg
