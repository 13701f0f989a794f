"""Full Sampler (RQSpline_MALA_Bundle: MALA local steps + flow training + NFProposal global steps) on the
GPU, with every strategy call checked against the oracle on the SAME inputs (teacher forcing: the
oracle receives the key, positions, buffers and flow parameters the GPU run had at that call).

Mirrors test/integration/test_quickstart.py and test/unit/test_bundle.py (runs to completion, repr)
and adds per-call value parity, which the reference's own tests do not pin.
"""
import numpy as np
import pytest
import torch

from flowutil import params_from_model
from parity import assert_close, compare_chains

pytestmark = pytest.mark.gpu


class Recorder:
    def __init__(self, name, inner, log, resources_snapshot):
        self.name, self.inner, self.log, self.snap = name, inner, log, resources_snapshot

    def __call__(self, rng_key, resources, x, data):
        entry = dict(name=self.name, key_in=np.array(rng_key, copy=True), x_in=x.detach().cpu().numpy().copy(),
                     pre=self.snap(self.name, resources, self.inner))
        out = self.inner(rng_key, resources, x, data)
        entry.update(key_out=np.array(out[0], copy=True), x_out=out[2].detach().cpu().numpy().copy(),
                     post=self.snap(self.name, out[1], self.inner))
        self.log.append(entry)
        return out


CONFIGS = {
    # a miniature of the quickstart that reaches the batched branch of sample_flow (n_global > n_NFproposal_batch_size)
    "small": dict(n_local=20, n_global=5, n_train=2, n_prod=2, n_epochs=2, hidden=[16, 16], n_layers=3, batch_size=128,
                  n_max_examples=400, nf_batch=3),
    # C1 = BASELINE.json configs[0] EXACTLY as SURVEY.md 8(d) states it (docs/tutorials/dualmoon.ipynb cells 5, 8, 11):
    # 5-D dual moon, 20 chains, 100 local + 10 global steps, 20 training + 20 production loops, 5 epochs, lr 5e-3,
    # batch_size = n_max_examples = 5000, MALA 0.1, flow 4 x [32, 32] x 8 bins, PRNGKey(42) split as in the notebook
    "C1": dict(n_local=100, n_global=10, n_train=20, n_prod=20, n_epochs=5, hidden=[32, 32], n_layers=4, batch_size=5000,
               n_max_examples=5000, nf_batch=10000),
}


def _moments_to_oracle(model, flat_dev):
    """A device-layout flat vector (Adam moment) in the oracle's flat order (no alignment padding)."""
    from oracle import nf
    mm = model.clone()
    mm.params.copy_(torch.from_numpy(np.asarray(flat_dev)).to(mm.params.device))
    return nf.flatten(params_from_model(mm))


@pytest.mark.parametrize("cfg_name", ["small", "C1"])
def test_quickstart_bundle_runs_and_every_call_matches_the_oracle(cuda, cfg_name):
    from flowmc_b200 import random as frandom, targets as T
    from flowmc_b200.Sampler import Sampler
    from flowmc_b200.resource_strategy_bundle.RQSpline_MALA import RQSpline_MALA_Bundle
    from oracle import local as olocal, nf, rng, targets as otargets

    c = CONFIGS[cfg_name]
    n_chains, d = 20, 5
    n_local, n_global, n_train, n_prod, n_epochs = c["n_local"], c["n_global"], c["n_train"], c["n_prod"], c["n_epochs"]
    key = frandom.PRNGKey(42)
    if cfg_name == "C1":      # notebook order: the bundle takes the first sub-key, the initial position the second
        key, sub = frandom.split(key)
        bundle_key = sub
        key, sub = frandom.split(key)
        x0 = frandom.normal(sub, (n_chains, d)) * 1
    else:
        key, sub = frandom.split(key)
        x0 = frandom.normal(sub, (n_chains, d))
        key, bundle_key = frandom.split(key)
    bundle = RQSpline_MALA_Bundle(bundle_key, n_chains, d, T.dual_moon(), n_local, n_global, n_train, n_prod, n_epochs,
                                  mala_step_size=0.1, rq_spline_hidden_units=c["hidden"], rq_spline_n_bins=8,
                                  rq_spline_n_layers=c["n_layers"], learning_rate=5e-3, batch_size=c["batch_size"],
                                  n_max_examples=c["n_max_examples"], n_NFproposal_batch_size=c["nf_batch"])
    assert repr(bundle) == "RQSpline_MALA Bundle"
    assert len(bundle.strategy_order) == 6 * n_train + 2 + 4 * n_prod

    log = []

    def snap(name, resources, strat):
        s = {}
        if name in ("local_stepper", "global_stepper"):
            st = resources["sampler_state"].data
            s["cursor"] = strat.current_position
            s["pos"] = resources[st["target_positions"]].data.cpu().numpy().copy()
            s["lp"] = resources[st["target_log_prob"]].data.cpu().numpy().copy()
            acc_name = st["target_local_accs"] if name == "local_stepper" else st["target_global_accs"]
            s["acc"] = resources[acc_name].data.cpu().numpy().copy()
        if name in ("global_stepper", "model_trainer"):
            s["flow"] = params_from_model(resources["model"] if name == "model_trainer"
                                          else resources["global_sampler"].model)
        if name == "model_trainer":
            s["buf"] = resources["positions_training"].data.cpu().numpy().copy()
            s["count"] = resources["optimizer"].optim_state.count
            s["mu"] = resources["optimizer"].optim_state.mu.cpu().numpy().copy()
            s["nu"] = resources["optimizer"].optim_state.nu.cpu().numpy().copy()
            s["loss"] = resources["loss_buffer"].data.cpu().numpy().copy()
        return s

    for nm in ("local_stepper", "global_stepper", "model_trainer"):
        bundle.strategies[nm] = Recorder(nm, bundle.strategies[nm], log, snap)
    sampler = Sampler(d, n_chains, key, resource_strategy_bundles=bundle)
    sampler.sample(x0, {})
    torch.cuda.synchronize()

    res = sampler.resources
    for nm in ("positions_training", "log_prob_training", "positions_production", "log_prob_production"):
        assert torch.isfinite(res[nm].data).all(), nm
    # local and global accept flags interleave at the shared cursor; unused slots stay -inf
    la, ga = res["local_accs_production"].data, res["global_accs_production"].data
    assert torch.isfinite(la[:, :n_local]).all() and torch.isinf(la[:, n_local:n_local + n_global]).all()
    assert torch.isinf(ga[:, :n_local]).all() and torch.isfinite(ga[:, n_local:n_local + n_global]).all()
    assert res["sampler_state"].data["training"] is False
    assert res["loss_buffer"].cursor == n_train * n_epochs and torch.isfinite(res["loss_buffer"].data).all()
    assert sampler.last_step.shape == (n_chains, d)

    packed = otargets.DualMoon.pack(d, None)
    mala = olocal.make_kernel("MALA", step_size=0.1)
    n_checked = {"local_stepper": 0, "global_stepper": 0, "model_trainer": 0}
    prev_key = None
    for e in log:
        if prev_key is not None:
            assert np.array_equal(e["key_in"], prev_key)      # keys thread through the strategy loop
        prev_key = e["key_out"]
        c0 = e["pre"].get("cursor", 0)
        if e["name"] == "local_stepper":
            o_key, o_pos, o_lp, o_acc, o_last, dbg = olocal.take_serial_steps(
                e["key_in"], e["x_in"], "dual_moon", packed, mala, n_local, return_debug=True)
            sl = slice(c0, c0 + n_local)
            assert np.array_equal(e["key_out"], o_key)
            compare_chains((e["post"]["pos"][:, sl], e["post"]["lp"][:, sl], e["post"]["acc"][:, sl]),
                           (o_pos, o_lp, o_acc), dbg, max_diverged_frac=0.1)
            assert e["post"]["cursor"] == c0 + n_local
        elif e["name"] == "global_stepper":
            o_key, o_pos, o_lp, o_acc, o_last, dbg = nf.take_group_steps(
                e["key_in"], e["x_in"], e["pre"]["flow"], "dual_moon", packed, n_global, c["nf_batch"])
            sl = slice(c0, c0 + n_global)
            assert np.array_equal(e["key_out"], o_key)
            compare_chains((e["post"]["pos"][:, sl], e["post"]["lp"][:, sl], e["post"]["acc"][:, sl]),
                           (o_pos, o_lp, o_acc), dbg["steps"], max_diverged_frac=0.1)
        else:
            k1, tkey, data, idx = nf.select_training_data(e["key_in"], e["pre"]["buf"], c["n_max_examples"], 100)
            ost = nf.AdamWState(nf.flatten(e["pre"]["flow"]).size)
            ost.count = e["pre"]["count"]
            if ost.count:
                # later trainings start from a warm optimiser state: the oracle gets the GPU run's Adam moments
                # (device layout = oracle layout + alignment padding; re-pack through the model's views)
                ost.mu = _moments_to_oracle(sampler.resources["model"], e["pre"]["mu"])
                ost.nu = _moments_to_oracle(sampler.resources["model"], e["pre"]["nu"])
            o_key, o_best, o_st, o_losses = nf.train(e["pre"]["flow"], tkey, data, ost, 5e-3, n_epochs, c["batch_size"])
            assert np.array_equal(e["key_out"], o_key)
            k0 = n_checked["model_trainer"] * n_epochs
            assert_close(e["post"]["loss"][k0:k0 + n_epochs], o_losses, "training losses", rtol=5e-4)
            q = e["post"]["flow"]
            assert_close(q.data_mean, o_best.data_mean, "data_mean", rtol=1e-4)
            for i in range(len(q.W)):
                assert_close(q.W[i], o_best.W[i], f"trained W[{i}]", rtol=5e-3)
            assert np.array_equal(e["x_in"], e["x_out"])
        n_checked[e["name"]] += 1
    assert n_checked["local_stepper"] == n_train + n_prod and n_checked["global_stepper"] == n_train + n_prod
    assert n_checked["model_trainer"] == n_train
    if cfg_name == "C1":      # the flow has learned the two moons by the production phase (dualmoon.ipynb's outcome)
        ga = res["global_accs_production"].data
        assert float(ga[torch.isfinite(ga)].mean()) > 0.05
        assert float(res["loss_buffer"].data[-1]) < float(res["loss_buffer"].data[0])


def test_sampler_dual_moon_statistics(cuda):
    """A slightly longer run: the flow learns the bimodal target and global acceptance becomes non-trivial."""
    from flowmc_b200 import random as frandom, targets as T
    from flowmc_b200.Sampler import Sampler
    from flowmc_b200.resource_strategy_bundle.RQSpline_MALA import RQSpline_MALA_Bundle
    n_chains, d = 64, 5
    key = frandom.PRNGKey(0)
    key, sub = frandom.split(key)
    x0 = frandom.normal(sub, (n_chains, d))
    key, sub = frandom.split(key)
    bundle = RQSpline_MALA_Bundle(sub, n_chains, d, T.dual_moon(), 50, 10, 6, 2, 5, mala_step_size=0.1,
                                  learning_rate=5e-3, batch_size=1000, n_max_examples=3000)
    s = Sampler(d, n_chains, key, resource_strategy_bundles=bundle)
    s.sample(x0, {})
    loss = s.resources["loss_buffer"].data.cpu().numpy()
    assert np.isfinite(loss).all() and loss[-1] < loss[0]
    ga = s.resources["global_accs_production"].data
    acc = ga[torch.isfinite(ga)].mean().item()
    assert acc > 0.02, f"global acceptance {acc}"
    pos = s.resources["positions_production"].data
    r = pos.norm(dim=-1)
    assert 1.0 < r.mean().item() < 4.0          # dual moon: mass near |x| = 2
