// jax.random-compatible counter RNG for sm_100a (threefry2x32-20, partitionable layout).
//
// Replaces the jax.random calls on the reference hot path (take_steps.py:71-72,158;
// MALA.py:62,66,83; HMC.py:128-136,144; Gaussian_random_walk.py:48-56; NF_proposal.py:41,99,103).
// Bit layout (jax >= 0.5.0, jax_threefry_partitionable=True): the element with row-major flat
// index i of any draw uses counter (hi32(i), lo32(i)); split() keeps both output words,
// random bits are o0 ^ o1.  Integer results are bit-exact by construction; the float
// transforms (uniform, normal via XLA's erf_inv polynomial) follow the same operation order.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace flowmc {

struct Key {
  uint32_t k0, k1;
};

__host__ __device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) {
#ifdef __CUDA_ARCH__
  return __funnelshift_l(x, x, r);
#else
  return (x << r) | (x >> (32 - r));
#endif
}

#define FLOWMC_TF_ROUND(r) \
  x0 += x1;                \
  x1 = rotl32(x1, r);      \
  x1 ^= x0;

// threefry2x32 with 20 rounds.  Key schedule constants are written so that, when the same key
// is used for many counters in an unrolled loop, the compiler hoists k2 and the k+i sums.
__host__ __device__ __forceinline__ void threefry2x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1,
                                                      uint32_t& o0, uint32_t& o1) {
  const uint32_t k2 = k0 ^ k1 ^ 0x1BD11BDAu;
  uint32_t x0 = c0 + k0;
  uint32_t x1 = c1 + k1;
  FLOWMC_TF_ROUND(13) FLOWMC_TF_ROUND(15) FLOWMC_TF_ROUND(26) FLOWMC_TF_ROUND(6)
  x0 += k1; x1 += k2 + 1u;
  FLOWMC_TF_ROUND(17) FLOWMC_TF_ROUND(29) FLOWMC_TF_ROUND(16) FLOWMC_TF_ROUND(24)
  x0 += k2; x1 += k0 + 2u;
  FLOWMC_TF_ROUND(13) FLOWMC_TF_ROUND(15) FLOWMC_TF_ROUND(26) FLOWMC_TF_ROUND(6)
  x0 += k0; x1 += k1 + 3u;
  FLOWMC_TF_ROUND(17) FLOWMC_TF_ROUND(29) FLOWMC_TF_ROUND(16) FLOWMC_TF_ROUND(24)
  x0 += k1; x1 += k2 + 4u;
  FLOWMC_TF_ROUND(13) FLOWMC_TF_ROUND(15) FLOWMC_TF_ROUND(26) FLOWMC_TF_ROUND(6)
  x0 += k2; x1 += k0 + 5u;
  o0 = x0;
  o1 = x1;
}

// split(key, n)[i]
__host__ __device__ __forceinline__ Key split_at(Key k, uint64_t i) {
  Key r;
  threefry2x32(k.k0, k.k1, (uint32_t)(i >> 32), (uint32_t)i, r.k0, r.k1);
  return r;
}

// random_bits(key, shape)[flat index i], 32-bit
__host__ __device__ __forceinline__ uint32_t bits_at(Key k, uint64_t i) {
  uint32_t a, b;
  threefry2x32(k.k0, k.k1, (uint32_t)(i >> 32), (uint32_t)i, a, b);
  return a ^ b;
}

#ifdef __CUDACC__
// [0,1) from the top 23 bits: bitcast((bits >> 9) | 0x3F800000) - 1
__device__ __forceinline__ float bits_to_unit(uint32_t bits) {
  return __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;
}

// jax.random.uniform(key) with minval=0, maxval=1
__device__ __forceinline__ float bits_to_uniform01(uint32_t bits) { return fmaxf(0.0f, bits_to_unit(bits)); }

// XLA ErfInv32 (Giles' single-precision polynomial): w = -log1p(-x*x); w < 5 ? P1(w - 2.5) : P2(sqrt(w) - 3);
// result p * x.  x*x is rounded on its own (as XLA's multiply op does) and log1p(y) is evaluated
// as ln2 * lg2(1 + y): 1 + y is exact for |x| >= 0.71 and otherwise off by < 3e-8, and w enters
// the polynomials only through (w - 2.5) / (sqrt(w) - 3), so the result agrees with the oracle's
// libm evaluation to ~1e-6 relative (tests/test_gpu_rng.py).  The three pieces are separate so
// that callers can run the branch-free central polynomial on every draw and patch the rare tail.
__device__ __forceinline__ float lg2_ftz(float x) {  // one MUFU.LG2; the argument is never denormal here
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// `valid == false` (a padding dimension) forces u = 0, hence w = 0 and a draw of exactly 0.
__device__ __forceinline__ float normal_arg(uint32_t bits, float& w, bool valid = true) {
  const float lo = -0.99999994f;  // nextafter(-1, 0): jax.random.normal's uniform minval
  float u = fmaxf(lo, fmaf(bits_to_unit(bits), 2.0f, lo));
  u = valid ? u : 0.0f;
  const float xx = __fmul_rn(u, u);
  w = -0.6931471805599453f * lg2_ftz(1.0f - xx);
  return u;
}
__device__ __forceinline__ float erf_inv_central(float w) {
  w -= 2.5f;
  float p = 2.81022636e-08f;
  p = fmaf(p, w, 3.43273939e-07f);
  p = fmaf(p, w, -3.5233877e-06f);
  p = fmaf(p, w, -4.39150654e-06f);
  p = fmaf(p, w, 0.00021858087f);
  p = fmaf(p, w, -0.00125372503f);
  p = fmaf(p, w, -0.00417768164f);
  p = fmaf(p, w, 0.246640727f);
  p = fmaf(p, w, 1.50140941f);
  return p;
}
__device__ __forceinline__ float erf_inv_tail(float w) {
  w = sqrtf(w) - 3.0f;
  float p = -0.000200214257f;
  p = fmaf(p, w, 0.000100950558f);
  p = fmaf(p, w, 0.00134934322f);
  p = fmaf(p, w, -0.00367342844f);
  p = fmaf(p, w, 0.00573950773f);
  p = fmaf(p, w, -0.0076224613f);
  p = fmaf(p, w, 0.00943887047f);
  p = fmaf(p, w, 1.00167406f);
  p = fmaf(p, w, 2.83297682f);
  return p;
}

// jax.random.normal: sqrt(2) * erf_inv(uniform(minval=nextafter(-1,0), maxval=1))
__device__ __forceinline__ float bits_to_normal(uint32_t bits) {
  float w;
  const float u = normal_arg(bits, w);
  const float p = (w < 5.0f) ? erf_inv_central(w) : erf_inv_tail(w);
  return 1.41421356237309515f * (p * u);
}
#endif

}  // namespace flowmc
