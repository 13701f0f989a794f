// dual_moon: docs/tutorials/dualmoon.ipynb:77-84 (mu = 0) and test/integration/test_MALA.py:14-23
// (mu = data["data"]):   logp = -( 0.5((|x-mu|-2)/0.1)^2 - lse(-0.5((x0 -/+ 3)/0.8)^2) - lse(-0.5((x1 -/+ 3)/0.6)^2) )
// data = mu[0..d).  Requires d >= 2.
#include "../../../include/flowmc_target.cuh"

struct DualMoon {
  static constexpr int NRED = 1;
  static constexpr bool USES_SCRATCH = false;
  struct Consts {
    const float* mu;
  };
  __device__ static Consts prepare(const float* data, int d) { return Consts{data}; }
  __device__ static float partial(const Consts& k, const flowmc::TargetCtx& c, int j, float xj, float* red) {
    const float r = xj - __ldg(k.mu + j);
    red[0] += r * r;
    return r;
  }
  __device__ static void lse_pair(float x, float w, float& lse, float& dlse) {
    const float a = (x - 3.0f) / w, b = (x + 3.0f) / w;
    const float ta = -0.5f * a * a, tb = -0.5f * b * b;
    const float m = fmaxf(ta, tb);
    const float ea = expf(ta - m), eb = expf(tb - m);
    const float s = ea + eb;
    lse = m + logf(s);
    dlse = (ea * (-a / w) + eb * (-b / w)) / s;
  }
  // red out: [0] = -(t/0.1)/|r|  (radial gradient factor)
  __device__ static float finish(const Consts& k, const flowmc::TargetCtx& c, float* red) {
    const float nrm = sqrtf(red[0]);
    const float t = (nrm - 2.0f) / 0.1f;
    float l2, d2, l3, d3;
    lse_pair(c.x[0], 0.8f, l2, d2);
    lse_pair(c.x[1], 0.6f, l3, d3);
    red[0] = -(t / 0.1f) / nrm;
    return -(0.5f * t * t - l2 - l3);
  }
  __device__ static float grad(const Consts& k, const flowmc::TargetCtx& c, int j, float xj, float aux,
                               const float* red) {
    float gj = red[0] * aux;
    if (j < 2) {
      float l, dl;
      lse_pair(xj, j == 0 ? 0.8f : 0.6f, l, dl);
      gj += dl;
    }
    return gj;
  }
};
FLOWMC_REGISTER_TARGET(DualMoon, "dual_moon")
