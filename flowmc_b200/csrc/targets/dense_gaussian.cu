// dense_gaussian (dense-precision variant of C2): logp = -1/2 x^T P x, grad = -P x, P symmetric.
// data = P row-major [d,d].
#include "../../../include/flowmc_target.cuh"

struct DenseGaussian {
  static constexpr int NRED = 1;
  static constexpr bool USES_SCRATCH = false;
  struct Consts {
    const float* P;
  };
  __device__ static Consts prepare(const float* data, int d) { return Consts{data}; }
  __device__ static float partial(const Consts& k, const flowmc::TargetCtx& c, int j, float xj, float* red) {
    const float* row = k.P + (int64_t)j * c.d;
    float s = 0.0f;
    for (int i = 0; i < c.d; ++i) s = fmaf(__ldg(row + i), c.x[i], s);
    red[0] += xj * s;
    return s;
  }
  __device__ static float finish(const Consts& k, const flowmc::TargetCtx& c, float* red) { return -0.5f * red[0]; }
  __device__ static float grad(const Consts& k, const flowmc::TargetCtx& c, int j, float xj, float aux,
                               const float* red) {
    return -aux;
  }
};
FLOWMC_REGISTER_TARGET(DenseGaussian, "dense_gaussian")
