// rosenbrock (BASELINE.json configs[2]):
//   logp = -sum_{i<d-1} [100 (x_{i+1} - x_i^2)^2 + (1 - x_i)^2] / 20.   No data.
#include "../../../include/flowmc_target.cuh"

struct Rosenbrock {
  static constexpr int NRED = 1;
  static constexpr bool USES_SCRATCH = false;
  struct Consts {};
  __device__ static Consts prepare(const float* data, int d) { return Consts{}; }
  __device__ static float partial(const Consts& k, const flowmc::TargetCtx& c, int j, float xj, float* red) {
    // branch-free: x[-1] and x[d] are readable (zero halo); boundary terms are masked by selects
    const float u = c.x[j + 1] - xj * xj;
    const float v = 1.0f - xj;
    const bool has_next = j < c.d - 1;
    // x / 20 as the exact 3-instruction constant division (flowmc::div_const): same bits as the division
    red[0] += has_next ? flowmc::div_const(100.0f * u * u + v * v, 20.0f, 0.05f) : 0.0f;
    float gj = has_next ? flowmc::div_const(400.0f * xj * u + 2.0f * v, 20.0f, 0.05f) : 0.0f;
    const float xm = c.x[j - 1];
    const float um = xj - xm * xm;
    gj += (j > 0) ? flowmc::div_const(-200.0f * um, 20.0f, 0.05f) : 0.0f;
    return gj;
  }
  __device__ static float finish(const Consts& k, const flowmc::TargetCtx& c, float* red) { return -red[0]; }
  __device__ static float grad(const Consts& k, const flowmc::TargetCtx& c, int j, float xj, float aux,
                               const float* red) {
    return aux;
  }
};
FLOWMC_REGISTER_TARGET(Rosenbrock, "rosenbrock")
