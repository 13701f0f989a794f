"""AdamOptimization on the GPU (flowmc_b200/strategy/optimization.py, csrc/adam_opt.cuh) vs the oracle restatement,
plus the reference's own tests for this strategy (test/unit/test_strategies.py:27-78) re-expressed on the device
path."""
import numpy as np
import pytest
import torch

from parity import assert_close

pytestmark = pytest.mark.gpu


def _strategy(**kw):
    from flowmc_b200 import targets as T
    from flowmc_b200.strategy.optimization import AdamOptimization
    args = dict(n_steps=100, learning_rate=5e-2, noise_level=0.0, bounds=np.array([[-np.inf, np.inf]]))
    args.update(kw)
    return AdamOptimization(T.iso_gaussian(0.5, "data"), **args)     # test_strategies.py:23-24


class TestOptimizationStrategies:
    n_dim = 2
    n_chains = 20
    n_steps = 100

    def _x0(self):
        from flowmc_b200 import random as frandom
        key = frandom.PRNGKey(42)
        key, subkey = frandom.split(key)
        return key, frandom.normal(subkey, (self.n_chains, self.n_dim)) * 1 + 10

    def test_repr(self, cuda):
        assert repr(_strategy()) == "AdamOptimization"

    def test_Adam_optimization(self, cuda):
        key, initial_position = self._x0()
        _, _, optimized_position = _strategy()(key, {}, initial_position, {"data": np.arange(self.n_dim)})
        assert optimized_position.shape == (self.n_chains, self.n_dim)
        assert torch.all(optimized_position.mean(dim=1) < initial_position.mean(dim=1))

    def test_standalone_optimize(self, cuda):
        key, initial_position = self._x0()
        _, optimized_position, final_log_prob = _strategy().optimize(key, None, initial_position,
                                                                     {"data": np.arange(self.n_dim)})
        assert optimized_position.shape == (self.n_chains, self.n_dim)
        assert torch.all(optimized_position.mean(dim=1) < initial_position.mean(dim=1))
        assert final_log_prob.shape == (self.n_chains,)
        assert torch.all(torch.isfinite(final_log_prob))

    def test_bounds_validation(self, cuda):
        from flowmc_b200 import targets as T
        from flowmc_b200.strategy.optimization import AdamOptimization
        with pytest.raises(ValueError, match="bounds must have shape"):
            AdamOptimization(T.iso_gaussian(0.5, "data"), bounds=np.array([-1.0, 1.0]))
        key, initial_position = self._x0()
        with pytest.raises(ValueError, match="incompatible with n_dim=2"):
            _strategy(bounds=np.zeros((3, 2)))(key, {}, initial_position, {"data": np.arange(self.n_dim)})
        with pytest.raises(TypeError):
            AdamOptimization(lambda x, data: -0.5 * (x ** 2).sum())


@pytest.mark.parametrize("d,n,steps,lr,noise,bounds", [
    (2, 20, 100, 5e-2, 0.0, [[-np.inf, np.inf]]),
    (5, 33, 40, 1e-2, 10.0, [[-np.inf, np.inf]]),             # the defaults' noisy gradient
    (12, 7, 25, 1e-1, 1.0, [[9.5, 10.5]]),                    # broadcast box: the projection is active
    (70, 5, 10, 3e-2, 0.5, None),                             # per-dimension box, d > 64 (third lane layout)
])
def test_matches_oracle(cuda, d, n, steps, lr, noise, bounds):
    from flowmc_b200 import random as frandom, targets as T
    from flowmc_b200.strategy.optimization import AdamOptimization
    from oracle import optimization as oopt, targets as O
    if bounds is None:
        bounds = np.stack([np.linspace(8.0, 9.9, d), np.linspace(10.1, 12.0, d)], axis=1)
    bounds = np.asarray(bounds, dtype=np.float32)
    mean = np.arange(d, dtype=np.float32) * 0.25
    key = frandom.PRNGKey(7)
    x0 = frandom.normal(frandom.PRNGKey(8), (n, d)) + 10
    strat = AdamOptimization(T.iso_gaussian(0.5, "data"), n_steps=steps, learning_rate=lr, noise_level=noise,
                             bounds=bounds)
    new_key, x, lp = strat.optimize(key, None, x0, {"data": mean})
    o_key, ox, olp = oopt.adam_optimize(key, "iso_gaussian", O.IsoGaussian.pack(d, 0.5, mean), x0.cpu().numpy(),
                                        n_steps=steps, learning_rate=lr, noise_level=noise, bounds=bounds)
    assert np.array_equal(new_key, o_key)
    # Adam divides by sqrt(nu): a 1-ulp difference in a tiny gradient is amplified early on, so free-running
    # trajectories are compared at the trajectory tolerance of the local kernels
    assert_close(x.cpu().numpy(), ox, "optimized positions", rtol=3e-4)
    assert_close(lp.cpu().numpy(), olp, "final log-prob", rtol=3e-4)
    assert float(x.min()) >= bounds[:, 0].min() and float(x.max()) <= bounds[:, 1].max()


def test_chain_sharding_is_bit_identical(cuda):
    from flowmc_b200 import random as frandom, targets as T
    from flowmc_b200.strategy.optimization import AdamOptimization
    d, n = 6, 40
    key = frandom.PRNGKey(3)
    x0 = frandom.normal(frandom.PRNGKey(4), (n, d)) + 3
    data = {"data": np.zeros(d, np.float32)}
    full = AdamOptimization(T.iso_gaussian(0.5, "data"), n_steps=30, noise_level=5.0)
    k_full, x_full, lp_full = full.optimize(key, None, x0, data)
    for a, b in ((0, 16), (16, 40)):
        part = AdamOptimization(T.iso_gaussian(0.5, "data"), n_steps=30, noise_level=5.0)
        part.set_chain_shard(a, n)
        k, x, lp = part.optimize(key, None, x0[a:b], data)
        assert np.array_equal(k, k_full)
        assert torch.equal(x, x_full[a:b]) and torch.equal(lp, lp_full[a:b])


def test_matches_the_committed_golden_vectors(cuda):
    """The device path against tests/golden/strategies_iso.npz (frozen oracle outputs on the reference's test set-up)."""
    import os
    from flowmc_b200 import random as frandom
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "strategies_iso.npz"))
    key = frandom.split(frandom.PRNGKey(42))[0]
    data = {"data": np.arange(2, dtype=np.float32)}
    k, x, lp = _strategy().optimize(key, None, torch.from_numpy(g["adam_x0"]).cuda(), data)
    assert np.array_equal(k, g["adam_key"])
    assert_close(x.cpu().numpy(), g["adam_x"], "adam positions", rtol=3e-4)
    assert_close(lp.cpu().numpy(), g["adam_lp"], "adam log-prob", rtol=3e-4)
    k, x, lp = _strategy(n_steps=30, learning_rate=1e-2, noise_level=10.0, bounds=np.array([[9.0, 10.5]])).optimize(
        key, None, torch.from_numpy(g["adam_x0"]).cuda(), data)
    assert_close(x.cpu().numpy(), g["adam_noisy_x"], "noisy adam positions", rtol=3e-4)
