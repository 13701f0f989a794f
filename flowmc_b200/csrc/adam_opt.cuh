// AdamOptimization (reference: src/flowMC/strategy/optimization.py:85-164): every chain runs n_steps of optax.adam
// on -logpdf with a noisy gradient and a box projection.  Same shape as the local-step kernels: one launch runs all
// steps for every chain, chain state (position, both Adam moments, key) in registers, the target's analytic
// gradient from the plugin header, jax.random-compatible keys:
//
//   rng_key, subkey = split(rng_key); keys = split(subkey, n_chains)                  (optimization.py:149-150)
//   per step:  key, sub = split(key)
//              grad = d(-logpdf)/dx * (1 + normal(sub) * noise_level)                 (:124-128)
//              mu = (1 - b1) g + b1 mu;  nu = (1 - b2) g^2 + b2 nu                     (optax.scale_by_adam)
//              x += -lr * (mu / (1 - b1^t)) / (sqrt(nu / (1 - b2^t)) + eps)
//              x = clip(x, lo, hi)                                                    (projection_box, :131-133)
//
// Every float operation is issued on its own (no FMA contraction), in optax's order; the bias corrections
// 1 - b^t come from the host as float32 (bc[t] = {1 - b1^(t+1), 1 - b2^(t+1)}).
#pragma once
#include "local_steps.cuh"
#include "registry.h"

namespace flowmc {

template <class T, class L>
__global__ void __launch_bounds__(32) adam_opt_kernel(const AdamOptArgs a) {
  constexpr int G = L::kG, DPL = L::kDPL, CPW = L::CPW, DS = L::DS;
  __shared__ __align__(16) float xrow_s[CPW][DS + 2 * kHalo];
  __shared__ __align__(16) float scratch[CPW][DS];
  const int lane = threadIdx.x, lg = lane % G, cw = lane / G;
  const int d = a.d;
  float* xrow = xrow_s[cw] + kHalo;
  if (lg == 0) {
#pragma unroll
    for (int q = 0; q < kHalo; ++q) {
      xrow[-1 - q] = 0.0f;
      xrow[DS + q] = 0.0f;
    }
  }
  __syncwarp();
  int64_t i = (int64_t)blockIdx.x * CPW + cw;
  const bool active = i < a.n_chains;
  if (!active) i = a.n_chains - 1;
  const typename T::Consts tc = T::prepare(a.data, d);
  float x[DPL], g[DPL], mu[DPL], nu[DPL], lo[DPL], hi[DPL];
#pragma unroll
  for (int k = 0; k < DPL; ++k) {
    const int j = L::dim(k, lg);
    const bool v = j < d;
    x[k] = v ? a.x0[i * d + j] : 0.0f;
    lo[k] = v ? a.lo[j] : 0.0f;
    hi[k] = v ? a.hi[j] : 0.0f;
    mu[k] = nu[k] = g[k] = 0.0f;
  }
  Key key = split_at(a.subkey, (uint64_t)(a.chain_offset + i));
  for (int t = 0; t < a.n_steps; ++t) {
    const Key sub = split_at(key, 1);
    key = split_at(key, 0);
    const float z = bits_to_normal(bits_at(sub, 0));                 // normal(sub), shape ()
    const float s = __fadd_rn(1.0f, __fmul_rn(z, a.noise_level));
    eval_target<T, L, true>(tc, x, g, xrow, scratch[cw], a.data, d, lg);
    const float bc1 = a.bc[2 * t], bc2 = a.bc[2 * t + 1];
#pragma unroll
    for (int k = 0; k < DPL; ++k) {
      const float gk = __fmul_rn(-g[k], s);                          // grad of -logpdf, noisy
      mu[k] = __fadd_rn(__fmul_rn(a.one_minus_b1, gk), __fmul_rn(a.b1, mu[k]));
      nu[k] = __fadd_rn(__fmul_rn(a.one_minus_b2, __fmul_rn(gk, gk)), __fmul_rn(a.b2, nu[k]));
      const float mh = __fdiv_rn(mu[k], bc1), nh = __fdiv_rn(nu[k], bc2);
      const float u = __fdiv_rn(mh, __fadd_rn(__fsqrt_rn(nh), a.eps));
      x[k] = __fadd_rn(x[k], __fmul_rn(a.neg_lr, u));
      x[k] = fminf(fmaxf(x[k], lo[k]), hi[k]);
    }
  }
  float lp = 0.0f;
  if (a.lp_out != nullptr) lp = eval_target<T, L, false>(tc, x, g, xrow, scratch[cw], a.data, d, lg);
  if (active) {
#pragma unroll
    for (int k = 0; k < DPL; ++k) {
      const int j = L::dim(k, lg);
      if (j < d) a.x_out[i * d + j] = x[k];
    }
    if (a.lp_out != nullptr && lg == 0) a.lp_out[i] = lp;
  }
}

template <class T, int G, int DPL, int VEC>
inline int launch_adam_one(const AdamOptArgs* a, cudaStream_t stream) {
  using L = Layout<G, DPL, VEC>;
  const int64_t nblk = (a->n_chains + L::CPW - 1) / L::CPW;
  adam_opt_kernel<T, L><<<(unsigned)nblk, 32, 0, stream>>>(*a);
  flowmc_count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    flowmc_set_error(cudaGetErrorString(e));
    return -3;
  }
  return 0;
}

template <class T>
int launch_adam_opt(const AdamOptArgs* a, cudaStream_t stream) {
  if (a->n_chains <= 0) return 0;
  if (a->d <= 8) return launch_adam_one<T, 1, 8, 1>(a, stream);
  if (a->d <= 64) return launch_adam_one<T, 8, 8, 1>(a, stream);
  if (a->d <= 512) return launch_adam_one<T, 32, 16, 1>(a, stream);
  flowmc_set_error("adam_optimize: unsupported dimension (d must be <= 512)");
  return -2;
}

}  // namespace flowmc
