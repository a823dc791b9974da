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

// split-bf16 ("fp32-grade") operands + similarity epilogues of the GEMM engine (gemm_sm100.cu / sim_loss.cu)
struct GemmSim {
  const void* a_lo;
  const void* b_lo;
  int epi;                   // 0 = plain GEMM over the three hi/lo products; 16 NCE fwd, 17 NCE bwd, 18 retrieval rank, 19 best match
  const float* temperature;
  int row_offset;
  void* part;                // float4 [M][part_ld]
  int part_ld;
  float* zt;
  const float* lse;
  float grad_scale;
  float* dtemp;
  int64_t lo_off;
  const float* best;
  const int32_t* bestj;
  const int64_t* lgid;
  const int64_t* rgid;
  int32_t* rank;
  unsigned long long* bestkey;
  const int32_t* tile_list;  // optional device list of mn tiles to visit (count in *tile_count); tile shape forced by the caller
  const int32_t* tile_count;
};
int gemm_sim_impl(Ctx*, const simseg_gemm_args*, const GemmSim*, cudaStream_t);

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

// ---- packed fp32x2 arithmetic (sm_100: FFMA2 / FMUL2 / FADD2 — one issue slot and one FMA-pipe pass for two lanes) ------
// The GELU epilogues are issue-bound: evaluating two adjacent columns per instruction halves the
// FMA-pipe instruction count.  A value is two fp32 in one 64-bit register pair (.x = low word).
// three-input maximum (FMNMX3 on sm_100)
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 f2_pack(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ f32x2 f2_splat(float a) { return f2_pack(a, a); }
__device__ __forceinline__ void f2_unpack(f32x2 p, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p));
}
__device__ __forceinline__ f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 f2_mul(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 f2_add(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// gelu of the two bf16 values in `u` -> packed fp32 pair.  One MUFU per value instead of two:
//   0.5*erfc(z) = 1 / (2^(1/16) * (1 + a1 z + ... + a6 z^6))^16,  z = |x|/sqrt2      (Abramowitz-Stegun 7.1.28, |err| <= 3e-7)
// evaluated in n = -|x| (sign bit OR-ed into the bf16 pair, odd coefficients negated) so that no negation is needed:
//   gelu(x) = max(x, 0) + n * w(n).      Absolute error of gelu <= 1e-6 (fp32 rounding through the 16th power included).
__device__ __forceinline__ f32x2 gelu_erf_x2(uint32_t u) {
  const float s = 1.04427378242741384f;                         // 2^(1/16): folds the 0.5 into the polynomial
  const float r2 = 0.70710678118654752f;
  const uint32_t un = u | 0x80008000u;
  const f32x2 n = f2_pack(bf16_lo(un), bf16_hi(un));            // -|x|
  f32x2 pl = f2_fma(f2_splat(s * 0.0000430638f * r2 * r2 * r2 * r2 * r2 * r2), n, f2_splat(-s * 0.0002765672f * r2 * r2 * r2 * r2 * r2));
  pl = f2_fma(pl, n, f2_splat(s * 0.0001520143f * r2 * r2 * r2 * r2));
  pl = f2_fma(pl, n, f2_splat(-s * 0.0092705272f * r2 * r2 * r2));
  pl = f2_fma(pl, n, f2_splat(s * 0.0422820123f * r2 * r2));
  pl = f2_fma(pl, n, f2_splat(-s * 0.0705230784f * r2));
  pl = f2_fma(pl, n, f2_splat(s));
  float p0, p1;
  f2_unpack(pl, p0, p1);
  f32x2 w = f2_pack(rcp_approx(p0), rcp_approx(p1));
  w = f2_mul(w, w);
  w = f2_mul(w, w);
  w = f2_mul(w, w);
  w = f2_mul(w, w);                                              // 0.5 * erfc(|x|/sqrt2)
  return f2_fma(n, w, f2_pack(fmaxf(bf16_lo(u), 0.f), fmaxf(bf16_hi(u), 0.f)));
}

// (gelu, gelu') of the two bf16 values in `u`: gelu_erf_both two columns per instruction.  Same arithmetic (Abramowitz-Stegun
// 7.1.26, exp(-x^2/2) shared by erfc and the density), same number of FMA-pipe lane operations, half the issue slots; the
// Phi(x) = x >= 0 ? 1 - w : w select stays scalar on the ALU pipe.
__device__ __forceinline__ void gelu_erf_both_x2(uint32_t u, f32x2& gelu, f32x2& dgelu) {
  const uint32_t ua = u & 0x7fff7fffu;
  const float x0 = bf16_lo(u), x1 = bf16_hi(u);
  const f32x2 x = f2_pack(x0, x1);
  const f32x2 ax = f2_pack(bf16_lo(ua), bf16_hi(ua));
  float t0, t1, e0, e1;
  f2_unpack(f2_fma(f2_splat(0.3275911f * 0.70710678118654752f), ax, f2_splat(1.0f)), t0, t1);
  f2_unpack(f2_mul(f2_mul(x, x), f2_splat(-0.5f * 1.44269504088896341f)), e0, e1);
  const f32x2 t = f2_pack(rcp_approx(t0), rcp_approx(t1));
  const f32x2 e = f2_pack(ex2_approx(e0), ex2_approx(e1));      // exp(-x^2/2)
  f32x2 q = f2_fma(f2_splat(0.5f * 1.061405429f), t, f2_splat(0.5f * -1.453152027f));
  q = f2_fma(q, t, f2_splat(0.5f * 1.421413741f));
  q = f2_fma(q, t, f2_splat(0.5f * -0.284496736f));
  q = f2_fma(q, t, f2_splat(0.5f * 0.254829592f));
  const f32x2 w = f2_mul(f2_mul(q, t), e);                      // 0.5 * erfc(|x|/sqrt2)
  const f32x2 omw = f2_fma(w, f2_splat(-1.0f), f2_splat(1.0f));
  float w0, w1, m0, m1;
  f2_unpack(w, w0, w1);
  f2_unpack(omw, m0, m1);
  const f32x2 cdf = f2_pack(x0 >= 0.f ? m0 : w0, x1 >= 0.f ? m1 : w1);
  gelu = f2_mul(x, cdf);
  dgelu = f2_fma(f2_mul(x, f2_splat(0.3989422804014327f)), e, cdf);
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
