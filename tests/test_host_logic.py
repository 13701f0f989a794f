"""Host-side logic that needs no GPU: jax.random-compatible key arithmetic through the C ABI, the temperature
adaptation and bias-correction tables of the 8f strategies, the prior family of TemperedPDF."""
import numpy as np
import pytest
import torch


def test_key_split_matches_the_oracle():
    from flowmc_b200 import random as frandom
    from oracle import rng
    key = frandom.PRNGKey(42)
    assert np.array_equal(frandom.split(key, 5), rng.split(rng.PRNGKey(42), 5))
    keys = frandom.split(frandom.PRNGKey(7), 9)
    got = frandom.split_each(keys, 4)                       # vmapped split: one C call
    want = np.stack([rng.split(k, 4) for k in keys])
    assert got.shape == (9, 4, 2) and np.array_equal(got, want)


def test_adapt_temperature_matches_the_oracle():
    from flowmc_b200.strategy.parallel_tempering import ParallelTempering
    from oracle import parallel_tempering as opt
    strat = ParallelTempering(4, "logpdf", "k", ["tp", "t"], "state")
    t = torch.arange(6, dtype=torch.float32) * 0.4 + 1
    acc = (torch.rand((64, 5), generator=torch.Generator().manual_seed(3)) < torch.tensor([0.9, 0.6, 0.5, 0.2, 0.7])).float()
    got = strat._adapt_temperature(t, acc).numpy()
    np.testing.assert_allclose(got, opt.adapt_temperature(t.numpy(), acc.numpy()), rtol=1e-6)
    assert got[0] == 1.0 and got[-1] == t[-1].item()        # the ladder's ends never move
    assert np.allclose(strat._adapt_temperature(t, torch.ones((7, 5))).numpy(), t.numpy())


def test_adam_bias_corrections():
    from flowmc_b200.strategy.optimization import adam_bias_corrections
    bc = adam_bias_corrections(5)
    assert bc.dtype == np.float32 and bc.shape == (5, 2)
    np.testing.assert_allclose(bc[:, 0], 1 - 0.9 ** np.arange(1, 6), rtol=1e-6)
    np.testing.assert_allclose(bc[:, 1], 1 - 0.999 ** np.arange(1, 6), rtol=2e-4)
    assert adam_bias_corrections(0).shape == (0, 2)


def test_box_quadratic_prior_family():
    from flowmc_b200.resource.logPDF import BoxQuadraticPrior
    from oracle import parallel_tempering as opt
    flat = BoxQuadraticPrior()
    assert flat.is_flat() and flat.packed(3).shape == (4, 3)
    x = torch.tensor([[0.5, -1.0, 2.0], [7.0, 0.0, 0.0]])
    assert torch.equal(flat(x), torch.zeros(2))
    pr = BoxQuadraticPrior(c=[0.5, 0.1, 0.0], mean=1.0, lower=-5.0, upper=5.0)
    assert not pr.is_flat()
    got = pr(x).numpy()
    want, _ = opt.log_prior(pr.packed(3), x.numpy())
    assert np.isneginf(got[1]) and np.isneginf(want[1])     # outside the box
    np.testing.assert_allclose(got[0], want[0], rtol=1e-6)


def test_adam_optimization_argument_checks_need_no_gpu():
    from flowmc_b200.strategy.optimization import AdamOptimization
    with pytest.raises(TypeError):
        AdamOptimization(lambda x, data: 0.0)


def test_nvtx_ranges_wrap_every_strategy_call():
    """flowmc_b200.tracing: one NVTX range per strategy call of Sampler.sample (no GPU needed: NVTX is a stub unless a
    profiler is attached)."""
    from flowmc_b200 import tracing
    from flowmc_b200.Sampler import Sampler
    calls = []

    def strat(name):
        def f(key, resources, x, data):
            calls.append(name)
            return key, resources, x
        return f
    s = Sampler(2, 1, np.zeros(2, np.uint32), resources={}, strategies={"a": strat("a"), "b": strat("b")},
                strategy_order=["a", "b", "a"])
    before = tracing.ranges_opened
    s.sample(np.zeros((1, 2), np.float32), {})
    assert calls == ["a", "b", "a"]
    assert tracing.ranges_opened - before in (0, 3)     # 3 when NVTX is loadable, 0 when tracing is unavailable
