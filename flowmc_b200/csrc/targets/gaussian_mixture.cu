// gaussian_mixture (BASELINE.json configs[3],[4]): K <= 8 isotropic components,
//   logp = logsumexp_k( logw_k - 0.5 * inv_var * |x - mu_k|^2 ),  grad = inv_var * sum_k softmax_k (mu_k - x).
// data = [K, inv_var, logw[0..K), mu[K][d]].
#include "../../../include/flowmc_target.cuh"

struct GaussianMixture {
  static constexpr int KMAX = 8;
  static constexpr int NRED = KMAX;
  static constexpr bool USES_SCRATCH = false;
  struct Consts {
    int K;
    float iv;
    const float* logw;
    const float* mu;
  };
  __device__ static Consts prepare(const float* data, int d) {
    const int K = (int)data[0];
    return Consts{K, data[1], data + 2, data + 2 + K};
  }
  __device__ static float partial(const Consts& k, const flowmc::TargetCtx& c, int j, float xj, float* red) {
#pragma unroll
    for (int m = 0; m < KMAX; ++m) {
      if (m < k.K) {
        const float r = __ldg(k.mu + (int64_t)m * c.d + j) - xj;
        red[m] += r * r;
      }
    }
    return 0.0f;
  }
  // red out: softmax weights
  __device__ static float finish(const Consts& k, const flowmc::TargetCtx& c, float* red) {
    float mx = -INFINITY;
#pragma unroll
    for (int m = 0; m < KMAX; ++m) {
      red[m] = (m < k.K) ? __ldg(k.logw + m) - 0.5f * k.iv * red[m] : -INFINITY;
      mx = fmaxf(mx, red[m]);
    }
    float s = 0.0f;
#pragma unroll
    for (int m = 0; m < KMAX; ++m) {
      red[m] = (m < k.K) ? expf(red[m] - mx) : 0.0f;
      s += red[m];
    }
    // e_m / s with one correctly rounded reciprocal + a residual correction each: the same bits as eight divisions
    const float rs = __frcp_rn(s);
#pragma unroll
    for (int m = 0; m < KMAX; ++m) red[m] = flowmc::div_const(red[m], s, rs);
    return mx + logf(s);
  }
  __device__ static float grad(const Consts& k, const flowmc::TargetCtx& c, int j, float xj, float aux,
                               const float* red) {
    float gj = 0.0f;
#pragma unroll
    for (int m = 0; m < KMAX; ++m)
      if (m < k.K) gj += red[m] * (__ldg(k.mu + (int64_t)m * c.d + j) - xj);
    return k.iv * gj;
  }
};
FLOWMC_REGISTER_TARGET(GaussianMixture, "gaussian_mixture")
