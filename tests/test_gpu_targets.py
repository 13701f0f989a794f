"""Device targets (logp + analytic gradient) vs the oracle's closed forms."""
import numpy as np
import pytest
import torch

from parity import assert_close

pytestmark = pytest.mark.gpu


QUARTIC_PLUGIN = r'''
#include "flowmc_target.cuh"
// logp = -c * sum x^4, data = [c]
struct Quartic {
  static constexpr int NRED = 1;
  static constexpr bool USES_SCRATCH = false;
  struct Consts { float c; };
  __device__ static Consts prepare(const float* data, int d) { return Consts{data[0]}; }
  __device__ static float partial(const Consts& k, const flowmc::TargetCtx& c, int j, float xj, float* red) {
    red[0] += k.c * xj * xj * xj * xj; return xj; }
  __device__ static float finish(const Consts& k, const flowmc::TargetCtx& c, float* red) { return -red[0]; }
  __device__ static float grad(const Consts& k, const flowmc::TargetCtx& c, int j, float xj, float aux,
                               const float* red) { return -4.0f * k.c * aux * aux * aux; }
};
FLOWMC_REGISTER_TARGET(Quartic, "test_quartic")
'''


def _cases():
    from flowmc_b200 import targets as T
    from oracle import targets as O
    rs = np.random.RandomState(0)
    P = rs.randn(24, 24)
    P = P @ P.T / 24 + np.eye(24)
    mu8 = rs.randn(8, 64) * 3
    return [
        ("iso_gaussian", T.iso_gaussian(0.5), {"data": np.arange(5, dtype=np.float32)}, 5,
         O.IsoGaussian.pack(5, 0.5, np.arange(5))),
        ("iso_gaussian", T.iso_gaussian(1.0, None), None, 3, O.IsoGaussian.pack(3, 1.0)),
        ("dual_moon", T.dual_moon(), None, 5, O.DualMoon.pack(5)),
        ("dual_moon", T.dual_moon("data"), {"data": np.arange(5)}, 5, O.DualMoon.pack(5, np.arange(5))),
        ("ar1_gaussian", T.ar1_gaussian(0.9), None, 128, O.AR1Gaussian.pack(128, 0.9)),
        ("ar1_gaussian", T.ar1_gaussian(0.5), None, 7, O.AR1Gaussian.pack(7, 0.5)),
        ("dense_gaussian", T.dense_gaussian(P), None, 24, O.DenseGaussian.pack(24, P)),
        ("rosenbrock", T.rosenbrock(), None, 64, O.Rosenbrock.pack(64)),
        ("rosenbrock", T.rosenbrock(), None, 6, O.Rosenbrock.pack(6)),
        ("gaussian_mixture", T.gaussian_mixture(mu8, 1.0), None, 64, O.GaussianMixture.pack(64, mu8, 1.0)),
        ("gaussian_mixture", T.gaussian_mixture(mu8[:3, :10], 0.5), None, 10,
         O.GaussianMixture.pack(10, mu8[:3, :10], 0.5)),
    ]


def test_logp_and_grad_match_oracle(cuda):
    from oracle import targets as O
    rs = np.random.RandomState(1)
    for name, tgt, data, d, packed in _cases():
        assert np.array_equal(tgt.pack(data, d), packed), name
        for n in (1, 33, 257):
            x = (rs.randn(n, d) * 1.5).astype(np.float32)
            xt = torch.from_numpy(x).to(cuda)
            lp, g = tgt.evaluate(xt, data, want_grad=True)
            lp_only = tgt.evaluate(xt, data)
            olp, og = O.logp_grad(name, x, packed)
            assert_close(lp.cpu().numpy(), olp, f"{name} d={d} logp", rtol=2e-5)
            assert torch.equal(lp, lp_only)
            assert_close(g.cpu().numpy(), og, f"{name} d={d} grad", rtol=2e-5)


def test_dual_moon_known_answer(cuda):
    # docs/tutorials/dualmoon.ipynb:65 -- the only known-answer value in the reference tree
    from flowmc_b200 import targets as T
    v = T.dual_moon().evaluate(torch.zeros(5, device=cuda), None)
    assert abs(float(v) - (-218.14496)) < 1e-3


def test_user_plugin_roundtrip(cuda, tmp_path):
    # a target written against include/flowmc_target.cuh, compiled and registered at run time
    from flowmc_b200 import targets as T
    src = QUARTIC_PLUGIN
    tgt = T.compile_target(src, "test_quartic", lambda data, d: np.array([0.25], np.float32), str(tmp_path))
    x = torch.randn(50, 12, device=cuda)
    lp, g = tgt.evaluate(x, None, want_grad=True)
    torch.testing.assert_close(lp, -(0.25 * x ** 4).sum(-1), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(g, -x ** 3, rtol=1e-5, atol=1e-5)
