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
__device__ __forceinline__ void gelu_erf_both(float x, float& gelu, float& dgelu) {
  const float z = fabsf(x) * 0.70710678118654752f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float e;                                             // exp(-z^2) = exp(-x^2/2)
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-z * z * 1.44269504088896341f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  const float erfc_half = 0.5f * poly * e;             // 0.5 * erfc(|x|/sqrt2)
  const float cdf = x >= 0.f ? 1.0f - erfc_half : erfc_half;
  gelu = x * cdf;
  dgelu = fmaf(x * 0.3989422804014327f, e, cdf);
}
__device__ __forceinline__ float gelu_erf(float x) {
  float g, d;
  gelu_erf_both(x, g, d);
  return g;
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
