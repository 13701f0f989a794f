// iso_gaussian: logp = -c * |x - mu|^2.  data = [c, mu[0..d)].
// Reference targets of this form: test/unit/test_kernels.py:14-15 (c=.5, mu=0),
// test/unit/test_strategies.py:23-24, test/integration/test_quickstart.py:7-8 (c=.5, mu=data["data"]),
// test/unit/test_resources.py:14-15 (c=1).
#include "../../../include/flowmc_target.cuh"

struct IsoGaussian {
  static constexpr int NRED = 1;
  static constexpr bool USES_SCRATCH = false;
  struct Consts {
    float c;
    const float* mu;
  };
  __device__ static Consts prepare(const float* data, int d) { return Consts{data[0], data + 1}; }
  __device__ static float partial(const Consts& k, const flowmc::TargetCtx& c, int j, float xj, float* red) {
    const float r = xj - __ldg(k.mu + j);
    red[0] += r * r;
    return r;
  }
  __device__ static float finish(const Consts& k, const flowmc::TargetCtx& c, float* red) { return -k.c * red[0]; }
  __device__ static float grad(const Consts& k, const flowmc::TargetCtx& c, int j, float xj, float aux,
                               const float* red) {
    return -2.0f * k.c * aux;
  }
};
FLOWMC_REGISTER_TARGET(IsoGaussian, "iso_gaussian")
