"""ParallelTempering on the GPU (flowmc_b200/strategy/parallel_tempering.py: tempered MALA steps in the persistent
local-step kernel, flowmc_pt_exchange, host temperature adaptation) vs the oracle restatement, plus the
reference's tests for this strategy (test/unit/test_strategies.py:337-520) re-expressed on the device path."""
import numpy as np
import pytest
import torch

from parity import assert_close

pytestmark = pytest.mark.gpu


def _prior_array(prior, d):
    return None if prior is None else prior.packed(d)


class TestTemperingStrategies:
    n_temps = 5
    n_dims = 3
    n_chains = 7
    n_steps = 4

    def initialize(self, prior=None, training=False, n_chains=None, n_dims=None, mala_step=1.0):
        from flowmc_b200 import random as frandom, targets as T
        from flowmc_b200.resource.buffers import Buffer
        from flowmc_b200.resource.kernel.MALA import MALA
        from flowmc_b200.resource.logPDF import TemperedPDF
        from flowmc_b200.resource.states import State
        from flowmc_b200.strategy.parallel_tempering import ParallelTempering
        n_chains = n_chains or self.n_chains
        n_dims = n_dims or self.n_dims
        mala = MALA(mala_step)
        logpdf = TemperedPDF(T.iso_gaussian(0.5, "data"), prior, n_dims=n_dims, n_temps=self.n_temps)
        key = frandom.PRNGKey(42)
        key, subkey = frandom.split(key)
        initial_position = frandom.normal(subkey, (n_chains, n_dims))
        key, subkey = frandom.split(key)
        tempered_initial_position = frandom.normal(subkey, (n_chains, self.n_temps - 1, n_dims))
        tempered_positions = Buffer("tempered_positions", (n_chains, self.n_temps - 1, n_dims), 2)
        tempered_positions.update_buffer(tempered_initial_position)
        temperatures = Buffer("temperatures", (self.n_temps,), 0)
        temperatures.update_buffer(torch.arange(self.n_temps) + 1.0)
        sampler_state = State({"target_positions": "tempered_positions", "target_log_prob": "logpdf",
                               "target_temperatures": "temperatures", "training": training}, name="sampler_state")
        resources = {"logpdf": logpdf, "MALA": mala, "tempered_positions": tempered_positions,
                     "temperatures": temperatures, "sampler_state": sampler_state}
        strat = ParallelTempering(n_steps=self.n_steps, tempered_logpdf_name="logpdf", kernel_name="MALA",
                                  tempered_buffer_names=["tempered_positions", "temperatures"],
                                  state_name="sampler_state")
        return key, resources, strat, initial_position

    def _data(self, d=None):
        return {"data": np.arange(d or self.n_dims, dtype=np.float32)}

    def test_tempered_log_pdf(self, cuda):
        key, resources, strat, x0 = self.initialize()
        logpdf = resources["logpdf"]
        base = logpdf(x0, self._data())
        t = torch.tensor(2.5, device=cuda)
        assert torch.allclose(logpdf.tempered_log_pdf(t, x0, self._data()), base / 2.5, rtol=1e-6)

    @pytest.mark.parametrize("with_prior", [False, True])
    def test_ensemble_step_matches_oracle(self, cuda, with_prior):
        from flowmc_b200 import random as frandom
        from flowmc_b200.resource.logPDF import BoxQuadraticPrior
        from oracle import parallel_tempering as opt, targets as O
        prior = BoxQuadraticPrior(c=0.05, mean=0.5, lower=-6.0, upper=6.0) if with_prior else None
        key, resources, strat, x0 = self.initialize(prior=prior, n_chains=33, n_dims=5, mala_step=0.7)
        strat.n_steps = 9
        positions = torch.cat([x0[:, None, :], resources["tempered_positions"].data], dim=1)
        temps = resources["temperatures"].data
        key, subkey = frandom.split(key)
        pos, lp, acc = strat._ensemble_steps(resources["MALA"], subkey, positions, resources["logpdf"], temps,
                                             self._data(5))
        o_pos, o_lp, o_acc = opt.ensemble_steps(subkey, positions.cpu().numpy(), "iso_gaussian",
                                                O.IsoGaussian.pack(5, 0.5, np.arange(5)), temps.cpu().numpy(), 9, 0.7,
                                                prior=_prior_array(prior, 5))
        assert pos.shape == (33, self.n_temps, 5) and lp.shape == (33, self.n_temps) and acc.shape == (33, self.n_temps, 9)
        same = (acc.cpu().numpy() == o_acc).all(axis=2)            # rows whose accept decisions all agree
        assert same.mean() > 0.9, "tempered MALA accept flags diverge from the oracle"
        assert_close(pos.cpu().numpy()[same], o_pos[same], "positions", rtol=3e-4)
        assert_close(lp.cpu().numpy()[same], o_lp[same], "tempered log-probs", rtol=3e-4)
        # hot rungs accept more: the ladder really tempers
        assert acc[:, -1].mean() >= acc[:, 0].mean() - 0.05

    def test_exchange_step_matches_oracle(self, cuda):
        from flowmc_b200 import random as frandom
        from oracle import parallel_tempering as opt, targets as O
        key, resources, strat, x0 = self.initialize(n_chains=200)
        positions = torch.cat([x0[:, None, :], resources["tempered_positions"].data], dim=1)
        temps = torch.arange(self.n_temps, device=cuda) * 0.3 + 1
        key, subkey = frandom.split(key)
        pos, lp, acc = strat._exchange(subkey, positions, resources["logpdf"], temps, self._data())
        o_pos, o_lp, o_acc, ratio, logu = opt.exchange(subkey, positions.cpu().numpy(), "iso_gaussian",
                                                       O.IsoGaussian.pack(3, 0.5, np.arange(3)), temps.cpu().numpy())
        assert acc.shape == (200, self.n_temps - 1)
        ga = acc.cpu().numpy()
        near = np.abs(ratio - logu) <= 1e-4 * np.maximum(1.0, np.abs(ratio))
        assert ((ga == o_acc) | near).all()
        same = (ga == o_acc).all(axis=1)
        assert same.mean() > 0.97
        assert np.array_equal(pos.cpu().numpy()[same], o_pos[same])        # swaps move rows, no arithmetic
        assert_close(lp.cpu().numpy()[same], o_lp[same], "exchanged log-probs")
        assert 0.05 < ga.mean() < 1.0

    def test_adapt_temperatures(self, cuda):
        from oracle import parallel_tempering as opt
        key, resources, strat, x0 = self.initialize()
        temperatures = torch.arange(self.n_temps, device=cuda) * 0.3 + 1
        out = strat._adapt_temperature(temperatures, torch.ones((self.n_chains, self.n_temps - 1), device=cuda))
        assert out.shape == (self.n_temps,)
        assert torch.allclose(out, temperatures)                     # equal acceptance everywhere: no change
        acc = (torch.rand((50, self.n_temps - 1), generator=torch.Generator().manual_seed(1)) < 0.5).float()
        got = strat._adapt_temperature(temperatures, acc.to(cuda)).cpu().numpy()
        np.testing.assert_allclose(got, opt.adapt_temperature(temperatures.cpu().numpy(), acc.numpy()), rtol=1e-6)

    @pytest.mark.parametrize("training", [False, True])
    def test_parallel_tempering(self, cuda, training):
        from oracle import parallel_tempering as opt, rng, targets as O
        key, resources, strat, x0 = self.initialize(training=training)
        before_pos = resources["tempered_positions"].data.clone()
        before_t = resources["temperatures"].data.clone()
        rng_key, resources, positions = strat(key, resources, x0, self._data())
        assert positions.shape == (self.n_chains, self.n_dims)
        o_key, o_p0, o_tp, o_t, _ = opt.parallel_tempering(key, x0.cpu().numpy(), before_pos.cpu().numpy(),
                                                           before_t.cpu().numpy(), "iso_gaussian",
                                                           O.IsoGaussian.pack(3, 0.5, np.arange(3)), self.n_steps, 1.0,
                                                           training=training)
        assert np.array_equal(rng_key, o_key)
        if not training:
            assert torch.equal(resources["tempered_positions"].data, before_pos)
            assert torch.equal(resources["temperatures"].data, before_t)
        else:
            assert not torch.equal(resources["tempered_positions"].data, before_pos)
            assert resources["temperatures"].data[0] == before_t[0] and resources["temperatures"].data[-1] == before_t[-1]
        # same chains as the oracle wherever no accept decision sat on a near-tie (7 chains x 5 rungs x 4 steps)
        close = np.isclose(positions.cpu().numpy(), o_p0, rtol=3e-4, atol=1e-5).all(axis=1)
        assert close.mean() >= 0.7

    def test_chain_sharding_is_bit_identical(self, cuda):
        key, resources, strat, x0 = self.initialize(n_chains=24)
        tp0 = resources["tempered_positions"].data.clone()
        _, _, full = strat(key, resources, x0, self._data())
        for a, b in ((0, 10), (10, 24)):
            k2, res2, strat2, x02 = self.initialize(n_chains=24)
            res2["tempered_positions"].data = tp0[a:b].clone()
            strat2.set_chain_shard(a, 24)
            _, _, part = strat2(k2, res2, x02[a:b], self._data())
            assert torch.equal(part, full[a:b])

    def test_only_local_kernels_and_device_priors(self, cuda):
        from flowmc_b200 import targets as T
        from flowmc_b200.resource.kernel.NF_proposal import NFProposal
        from flowmc_b200.resource.logPDF import TemperedPDF
        from flowmc_b200.resource.model.nf_model.rqSpline import MaskedCouplingRQSpline
        with pytest.raises(TypeError):
            TemperedPDF(T.iso_gaussian(0.5, "data"), lambda x, data: 0.0, n_dims=3)
        key, resources, strat, x0 = self.initialize()
        resources["MALA"] = NFProposal(MaskedCouplingRQSpline(3, 2, [16, 16], 8, key))
        with pytest.raises(NotImplementedError):
            strat(key, resources, x0, self._data())

    @pytest.mark.parametrize("kind", ["HMC", "GRW"])
    @pytest.mark.parametrize("with_prior", [False, True])
    def test_any_local_kernel_as_the_tempered_kernel(self, cuda, kind, with_prior):
        """The reference hands whatever ProposalBase the resources name to _individual_step
        (parallel_tempering.py:74,135-289): HMC (dense inverse mass) and the Gaussian random walk on the tempered
        density, against the oracle -- ensemble steps and one full ParallelTempering call."""
        from flowmc_b200 import random as frandom
        from flowmc_b200.resource.kernel.Gaussian_random_walk import GaussianRandomWalk
        from flowmc_b200.resource.kernel.HMC import HMC
        from flowmc_b200.resource.logPDF import BoxQuadraticPrior
        from oracle import parallel_tempering as opt, targets as O
        d = 5
        prior = BoxQuadraticPrior(c=0.05, mean=0.5, lower=-6.0, upper=6.0) if with_prior else None
        key, resources, strat, x0 = self.initialize(prior=prior, n_chains=33, n_dims=d, training=True)
        if kind == "HMC":
            rs = np.random.RandomState(2)
            A = rs.randn(d, d) * 0.2
            M = (A @ A.T + np.eye(d)).astype(np.float32)
            resources["MALA"] = HMC(M, 0.3, 3)
            okw = dict(kind="HMC", n_leapfrog=3, condition_matrix=M)
            step = 0.3
        else:
            resources["MALA"] = GaussianRandomWalk(0.6)
            okw = dict(kind="GRW")
            step = 0.6
        strat.n_steps = 7
        positions = torch.cat([x0[:, None, :], resources["tempered_positions"].data], dim=1)
        temps = resources["temperatures"].data
        k1, subkey = frandom.split(key)
        pos, lp, acc = strat._ensemble_steps(resources["MALA"], subkey, positions, resources["logpdf"], temps,
                                             self._data(d))
        packed = O.IsoGaussian.pack(d, 0.5, np.arange(d))
        o_pos, o_lp, o_acc = opt.ensemble_steps(subkey, positions.cpu().numpy(), "iso_gaussian", packed,
                                                temps.cpu().numpy(), 7, step, prior=_prior_array(prior, d), **okw)
        same = (acc.cpu().numpy() == o_acc).all(axis=2)
        assert same.mean() > 0.9, f"tempered {kind} accept flags diverge from the oracle"
        assert 0.05 < float(acc.mean()) < 1.0
        assert_close(pos.cpu().numpy()[same], o_pos[same], "positions", rtol=3e-4)
        assert_close(lp.cpu().numpy()[same], o_lp[same], "tempered log-probs", rtol=3e-4)
        # one full call: keys, cold-chain positions, adapted ladder
        tp0 = resources["tempered_positions"].data.cpu().numpy().copy()
        t0 = temps.cpu().numpy().copy()               # (the training call adapts the ladder in place)
        new_key, res, cold = strat(key, resources, x0, self._data(d))
        o_key, o_cold, o_tp, o_t, o_acc2 = opt.parallel_tempering(key, x0.cpu().numpy(), tp0, t0,
                                                                  "iso_gaussian", packed, 7, step,
                                                                  prior=_prior_array(prior, d), training=True, **okw)
        assert np.array_equal(new_key, o_key)
        agree = np.abs(cold.cpu().numpy() - o_cold).max(axis=1) <= 3e-4 * max(1.0, float(np.abs(o_cold).max()))
        assert agree.mean() >= 0.7          # chains whose accept / swap decisions all agree land on the same point
        np.testing.assert_allclose(res["temperatures"].data.cpu().numpy(), o_t, rtol=0.2)


def test_rqspline_mala_pt_bundle(cuda):
    """test/unit/test_bundle.py:35-58 (construction, repr) + a short end-to-end Sampler run through the bundle."""
    from flowmc_b200 import random as frandom, targets as T
    from flowmc_b200.Sampler import Sampler
    from flowmc_b200.resource.logPDF import BoxQuadraticPrior
    from flowmc_b200.resource_strategy_bundle.RQSpline_MALA_PT import RQSpline_MALA_PT_Bundle
    n_chains, n_dims = 16, 3
    rng_key = frandom.PRNGKey(0)
    bundle = RQSpline_MALA_PT_Bundle(rng_key=rng_key, n_chains=n_chains, n_dims=n_dims, logpdf=T.iso_gaussian(0.5, "data"),
                                     n_local_steps=10, n_global_steps=5, n_training_loops=2, n_production_loops=1,
                                     n_epochs=3, batch_size=64, n_max_examples=200,
                                     logprior=BoxQuadraticPrior(c=0.01, lower=-20.0, upper=20.0))
    assert repr(bundle) == "RQSpline MALA PT Bundle"
    assert bundle.strategy_order[0] == "initialize_tempered_positions"
    assert bundle.strategy_order.count("parallel_tempering") == 3
    assert bundle.resources["tempered_positions"].shape == (n_chains, 4, n_dims)
    assert torch.allclose(bundle.resources["temperatures"].data.cpu(), torch.linspace(1.0, 5.0, 5))
    key, sub = frandom.split(frandom.PRNGKey(1))
    sampler = Sampler(n_dims, n_chains, key, resource_strategy_bundles=bundle)
    x0 = frandom.normal(sub, (n_chains, n_dims))
    sampler.sample(x0, {"data": np.arange(n_dims, dtype=np.float32)})
    prod = sampler.resources["positions_production"].data
    assert prod.shape == (n_chains, 15, n_dims) and bool(torch.isfinite(prod).all())
    tp = sampler.resources["tempered_positions"].data
    assert bool(torch.isfinite(tp).all()) and not torch.equal(tp[:, 0], tp[:, 1])
    t = sampler.resources["temperatures"].data
    assert float(t[0]) == 1.0 and float(t[-1]) == 5.0 and bool(torch.isfinite(t).all())
    assert sampler.resources["sampler_state"].data["training"] is False
