// Counter-based random bits for train-mode dropout (HF BertModel: hidden_dropout_prob / attention_probs_dropout_prob,
// active under model.train()).  Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as
// easy as 1, 2, 3", SC'11): a mask bit is a pure function of (seed, step, site, element), so backward regenerates exactly
// the mask forward used and nothing but the two 64-bit words {seed, step} is kept.  Those live in DEVICE memory and are read
// by the kernels — a replayed CUDA graph therefore draws a fresh mask every step (the host bumps `step` with a captured add).
//   key     = {seed lo, seed hi}
//   counter = {element group, row, site, step lo}       one call -> four 32-bit words -> four consecutive elements
//   element kept  <=>  word >= thr,  thr = round(p * 2^32);   kept values are scaled by 1 / (1 - p)
// oracle/simseg_oracle.py:philox4x32_10 is the numpy restatement the tests compare against bit for bit.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace simseg {

struct DropSpec {
  const unsigned long long* rng;   // device {seed, step}
  uint32_t site;
  uint32_t thr;                    // drop iff word < thr
  float inv_keep;                  // 1 / (1 - p)
};

__host__ inline DropSpec make_drop_spec(float p, const void* rng, uint32_t site) {
  DropSpec d;
  d.rng = reinterpret_cast<const unsigned long long*>(rng);
  d.site = site;
  const double t = static_cast<double>(p) * 4294967296.0;
  d.thr = t >= 4294967295.0 ? 0xffffffffu : static_cast<uint32_t>(t + 0.5);
  d.inv_keep = 1.0f / (1.0f - p);
  return d;
}

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

struct DropKey {
  uint2 key;
  uint32_t step;
};
__device__ __forceinline__ DropKey load_drop_key(const DropSpec& d) {
  const unsigned long long seed = d.rng[0], step = d.rng[1];
  DropKey k;
  k.key = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  k.step = static_cast<uint32_t>(step);
  return k;
}

// the four words of elements 4*group .. 4*group+3 of `row`
__device__ __forceinline__ uint4 drop_words(const DropSpec& d, const DropKey& k, uint32_t group, uint32_t row) {
  return philox4x32_10(make_uint4(group, row, d.site, k.step), k.key);
}

}  // namespace simseg
