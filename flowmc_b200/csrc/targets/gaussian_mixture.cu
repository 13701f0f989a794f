// gaussian_mixture (BASELINE.json configs[3],[4]): K <= 8 isotropic components,
//   logp = logsumexp_k( logw_k - 0.5 * inv_var * |x - mu_k|^2 ),  grad = inv_var * sum_k softmax_k (mu_k - x).
// data = [K, inv_var, logw[0..K), mu[K][d]].
#include "../../../include/flowmc_target.cuh"

struct GaussianMixture {
  static constexpr int KMAX = 8;
  static constexpr int NRED = KMAX;
  static constexpr bool USES_SCRATCH = false;
  __device__ static float partial(const flowmc::TargetCtx& c, int j, float xj, float* red) {
    const int K = (int)c.data[0];
    const float* mu = c.data + 2 + K + j;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      if (k < K) {
        const float r = __ldg(mu + (int64_t)k * c.d) - xj;
        red[k] += r * r;
      }
    }
    return 0.0f;
  }
  // red out: softmax weights
  __device__ static float finish(const flowmc::TargetCtx& c, float* red) {
    const int K = (int)c.data[0];
    const float iv = c.data[1];
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      red[k] = (k < K) ? c.data[2 + k] - 0.5f * iv * red[k] : -INFINITY;
      m = fmaxf(m, red[k]);
    }
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      red[k] = (k < K) ? expf(red[k] - m) : 0.0f;
      s += red[k];
    }
#pragma unroll
    for (int k = 0; k < KMAX; ++k) red[k] = red[k] / s;
    return m + logf(s);
  }
  __device__ static float grad(const flowmc::TargetCtx& c, int j, float xj, float aux, const float* red) {
    const int K = (int)c.data[0];
    const float* mu = c.data + 2 + K + j;
    float gj = 0.0f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < K) gj += red[k] * (__ldg(mu + (int64_t)k * c.d) - xj);
    return c.data[1] * gj;
  }
};
FLOWMC_REGISTER_TARGET(GaussianMixture, "gaussian_mixture")
