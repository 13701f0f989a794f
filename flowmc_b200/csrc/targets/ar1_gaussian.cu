// ar1_gaussian (BASELINE.json configs[1]): zero-mean Gaussian with covariance
// rho^|i-j|, i.e. tridiagonal precision P = 1/(1-rho^2) * tridiag(-rho, 1+rho^2 (1 at the ends), -rho).
// logp = -1/2 x^T P x,  grad = -P x.   data = [rho].
#include "../../../include/flowmc_target.cuh"

struct AR1Gaussian {
  static constexpr int NRED = 1;
  static constexpr bool USES_SCRATCH = false;
  struct Consts {
    float rho, a, dmid;
  };
  __device__ static Consts prepare(const float* data, int d) {
    const float rho = data[0];
    return Consts{rho, 1.0f / (1.0f - rho * rho), 1.0f + rho * rho};
  }
  __device__ static float partial(const Consts& k, const flowmc::TargetCtx& c, int j, float xj, float* red) {
    // x[-1] and x[d] are zero (halo guaranteed by the kernels), so no bounds tests on the neighbours
    const float diag = (j > 0 && j < c.d - 1) ? k.dmid : 1.0f;
    const float px = k.a * (diag * xj - k.rho * (c.x[j - 1] + c.x[j + 1]));
    red[0] += xj * px;
    return px;
  }
  __device__ static float finish(const Consts& k, const flowmc::TargetCtx& c, float* red) { return -0.5f * red[0]; }
  __device__ static float grad(const Consts& k, const flowmc::TargetCtx& c, int j, float xj, float aux,
                               const float* red) {
    return -aux;
  }
};
FLOWMC_REGISTER_TARGET(AR1Gaussian, "ar1_gaussian")
