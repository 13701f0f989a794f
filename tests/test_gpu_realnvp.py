"""RealNVP on the GPU (csrc/flow_realnvp.cu through flowmc_b200.resource.model.nf_model.realNVP) vs the oracle
(oracle/realnvp.py) and the committed golden fixture: bijection, log_prob, sample, bit-level initialisation, the
hand-written loss gradient against float64 autograd, a short training trajectory, NFProposal global steps with a
RealNVP proposal, and the reference's own tests for this model (test/unit/test_nf.py:9-52,
test/integration/test_normalizingFlow.py:10-28) re-expressed.

Reference: src/flowMC/resource/model/nf_model/realNVP.py:18-228, resource/model/common.py:68-209."""
import os
import sys

import numpy as np
import pytest
import torch

from flowutil import nvp_model_from_params, nvp_params_from_model
from parity import assert_close

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
G = np.load(os.path.join(HERE, "golden", "realnvp_d5.npz"))

SHAPES = [(5, 4, 16, 24), (2, 4, 32, 100), (3, 2, 4, 2), (32, 10, 128, 300), (64, 6, 100, 77), (7, 3, 33, 65)]


def _params(seed, d, L, h):
    from make_golden import realnvp_params
    return realnvp_params(seed, d, L, h)


def _flat_to_oracle(p, m, flat):
    mm = m.clone()
    mm.params.copy_(flat)
    return nvp_params_from_model(mm)


def test_golden_fixture(cuda):
    """Fixed known-answer inputs / outputs (tests/golden/realnvp_d5.npz, written by tests/golden/make_golden.py)."""
    d, L, h = [int(v) for v in G["shape"]]
    p = _params(31, d, L, h)
    m = nvp_model_from_params(p)
    x = torch.from_numpy(G["x"]).cuda()
    y, ld = m.forward(x)
    assert_close(y.cpu().numpy(), G["fwd_y"], "forward y")
    assert_close(ld.cpu().numpy(), G["fwd_logdet"], "forward logdet")
    xi, ldi = m.inverse(x)
    assert_close(xi.cpu().numpy(), G["inv_x"], "inverse x")
    assert_close(ldi.cpu().numpy(), G["inv_logdet"], "inverse logdet")
    assert_close(m.log_prob(x).cpu().numpy(), G["log_prob"], "log_prob")
    assert_close(m.sample(G["sample_key"], 16).cpu().numpy(), G["sample"], "sample")
    loss, grad = m.loss_and_grad(x)
    assert abs(float(loss.item()) - float(G["loss"])) <= 1e-5 * max(1.0, abs(float(G["loss"])))
    from oracle import realnvp as onvp
    g = onvp.flatten(_flat_to_oracle(p, m, grad))
    ref = G["grad_flat"]
    assert np.abs(g - ref).max() <= 5e-5 * np.abs(ref).max()


@pytest.mark.parametrize("d,L,h,n", SHAPES)
def test_bijection_log_prob_sample_match_oracle(cuda, d, L, h, n):
    from flowmc_b200 import random as frandom
    from oracle import realnvp as onvp
    p = _params(100 + d, d, L, h)
    m = nvp_model_from_params(p)
    x = (1.5 * np.random.default_rng(d).standard_normal((n, d))).astype(np.float32)
    xd = torch.from_numpy(x).cuda()
    y, ld = m.forward(xd)
    oy, old = onvp.forward(p, x)
    assert_close(y.cpu().numpy(), oy, "forward")
    assert_close(ld.cpu().numpy(), old, "forward logdet")
    xi, ldi = m.inverse(xd)
    ox, oldi = onvp.inverse(p, x)
    assert_close(xi.cpu().numpy(), ox, "inverse")
    assert_close(ldi.cpu().numpy(), oldi, "inverse logdet")
    assert_close(m.log_prob(xd).cpu().numpy(), onvp.log_prob(p, x), "log_prob")
    key = frandom.PRNGKey(d + 1)
    assert_close(m.sample(key, n).cpu().numpy(), onvp.sample(p, key, n), "sample")
    # single-sample form of the API
    y1, ld1 = m.forward(xd[0])
    assert y1.shape == (d,) and ld1.dim() == 0 and torch.equal(y1, y[0])
    assert m.log_prob(xd[0]).dim() == 0


def test_initialisation_is_bit_exact(cuda):
    from flowmc_b200 import random as frandom
    from flowmc_b200.resource.model.nf_model.realNVP import RealNVP
    from oracle import realnvp as onvp
    for seed, d, L, h in ((0, 3, 2, 4), (42, 5, 4, 32), (7, 32, 3, 128)):
        m = RealNVP(d, L, h, frandom.PRNGKey(seed))
        q = nvp_params_from_model(m)
        p = onvp.init_params(frandom.PRNGKey(seed), d, L, h)
        assert np.array_equal(q.mask, p.mask)
        for name in ("b1", "W2", "b2"):          # uniform draws: same RNG words, same float32 transform
            np.testing.assert_allclose(getattr(q, name), getattr(p, name), rtol=3e-6, atol=1e-8, err_msg=name)
        np.testing.assert_allclose(q.W1, p.W1, rtol=3e-6, atol=1e-9)   # normal draws x sqrt(1e-4 / d)
    m = RealNVP(3, 2, 4, frandom.PRNGKey(0))
    x = torch.tensor([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]]).cuda()
    assert_close(m.log_prob(x).cpu().numpy(), G["init_key0_log_prob"], "log_prob of the key-0 model")


def test_reference_test_nf_realnvp_and_affine_coupling(cuda):
    """test/unit/test_nf.py:9-52."""
    from flowmc_b200 import random as frandom
    from flowmc_b200.resource.model.nf_model.realNVP import AffineCoupling, RealNVP
    # test_affine_coupling_forward_and_inverse
    x = torch.tensor([[1.0, 2.0], [3.0, 4.0]]).cuda()
    mask = np.where(np.arange(2) % 2 == 0, 1.0, 0.0)
    layer = AffineCoupling(2, 4, mask, frandom.PRNGKey(0), 0.5)
    y, ld = layer.forward(x)
    xr, ldi = layer.inverse(y)
    assert torch.allclose(x, torch.round(xr, decimals=5))
    assert torch.allclose(ld, -ldi)
    assert torch.equal(y[:, 1], x[:, 1]) and not torch.equal(y[:, 0], x[:, 0])   # mask = 1 entries are transformed
    # test_realnvp
    rng_key, _ = frandom.split(frandom.PRNGKey(0), 2)
    model = RealNVP(3, 2, 4, rng_key)
    x = torch.tensor([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]]).cuda()
    y, log_det = model(x)
    assert y.shape == x.shape and log_det.shape == (2,)
    y_inv, log_det_inv = model.inverse(y)
    assert y_inv.shape == x.shape and log_det_inv.shape == (2,)
    assert torch.allclose(x, y_inv) and torch.allclose(log_det, -log_det_inv)
    samples = model.sample(frandom.PRNGKey(0), 2)
    assert samples.shape == (2, 3) and model.log_prob(samples).shape == (2,)
    assert repr(model) == "RealNVP with n_features=3, n_layers=2"


@pytest.mark.parametrize("d,L,h,n", [(5, 4, 16, 100), (32, 10, 128, 260), (3, 2, 4, 33), (64, 4, 96, 70)])
def test_loss_and_grad_match_float64_autograd(cuda, d, L, h, n):
    from oracle import realnvp as onvp
    p = _params(17 + d, d, L, h)
    m = nvp_model_from_params(p)
    x = (1.5 * np.random.default_rng(1).standard_normal((n, d))).astype(np.float32)
    loss, grad = m.loss_and_grad(torch.from_numpy(x).cuda())
    o_loss, og = onvp.loss_and_grads(p, x)
    assert abs(float(loss.item()) - o_loss) <= 1e-5 * max(1.0, abs(o_loss))
    q = _flat_to_oracle(p, m, grad)
    for name in ("W1", "b1", "W2", "b2"):
        ref = og[name]
        err = np.abs(getattr(q, name) - ref).max()
        tol = 5e-5 * np.abs(ref).max() + 1e-7
        assert err <= tol, f"d{name}: max err {err:.3e} > {tol:.3e}"
    # masks, whitening constants and the base distribution get exactly zero gradient
    assert not q.mask.any() and not q.data_mean.any() and not q.data_cov.any() and not q.base_cov.any()
    # row gather + data-parallel slices: half-batch gradients scaled by 1 / n_total add up to the full batch's
    xd = torch.from_numpy(x).cuda()
    idx = torch.arange(n, dtype=torch.int32, device="cuda")
    l_full, g_full = [t.clone() for t in m.loss_and_grad(xd, idx)]
    la, ga = [t.clone() for t in m.loss_and_grad(xd, idx[: n // 2].contiguous(), n_global=n)]
    lb, gb = [t.clone() for t in m.loss_and_grad(xd, idx[n // 2:].contiguous(), n_global=n)]
    assert abs(float(la + lb) - float(l_full)) <= 1e-5 * max(1.0, abs(float(l_full)))
    assert float((ga + gb - g_full).abs().max()) <= 2e-5 * float(g_full.abs().max())


def test_train_matches_oracle_for_a_few_steps(cuda):
    """NFModel.train with a RealNVP: data statistics, epoch key schedule, permutation batches, fused clip + AdamW over
    the flat blob -- including the weight decay of the float masks."""
    from flowmc_b200.resource.model.nf_model.realNVP import RealNVP
    from flowmc_b200.resource.optimizer import Optimizer
    from oracle import nf, realnvp as onvp, rng
    d, L, h = 2, 4, 32
    key = rng.PRNGKey(2)
    p = onvp.init_params(key, d, L, h)
    m = nvp_model_from_params(p)
    z = rng.normal(rng.PRNGKey(4), (700, 2))
    data = np.stack([z[:, 0], z[:, 0] ** 2 + np.float32(0.3) * z[:, 1]], axis=1).astype(np.float32)
    opt = Optimizer(m, learning_rate=5e-3)
    tkey = rng.PRNGKey(4)
    out_key, best, best_state, losses = m.train(tkey, torch.from_numpy(data).cuda(), opt.optim, opt.optim_state,
                                                num_epochs=3, batch_size=256, verbose=False)
    o_key, o_best, o_state, o_losses = onvp.train(p, tkey, data, nf.AdamWState(onvp.flatten(p).size), 5e-3, 3, 256)
    assert np.array_equal(out_key, o_key)
    assert_close(losses.cpu().numpy(), o_losses, "epoch losses", rtol=2e-4)
    q = nvp_params_from_model(best)
    for name in ("W1", "b1", "W2", "b2"):
        assert_close(getattr(q, name), getattr(o_best, name), f"{name} after training", rtol=5e-4)
    assert_close(q.mask, o_best.mask, "weight-decayed masks", rtol=1e-6)
    assert float(q.mask.max()) < 1.0
    assert best_state.count == o_state.count
    assert isinstance(best, RealNVP) and best is not m
    assert torch.equal(m.params, nvp_model_from_params(p).params) and opt.optim_state.count == 0


def test_reference_integration_test_realnvp(cuda):
    """test/integration/test_normalizingFlow.py:10-28: RealNVP(2, 4, 32), 5 epochs of batch 100, then 10000 samples."""
    from flowmc_b200 import random as frandom
    from flowmc_b200.resource.model.nf_model.realNVP import RealNVP
    from flowmc_b200.resource.optimizer import Optimizer
    key1, rng_, init_rng = frandom.split(frandom.PRNGKey(0), 3)
    data = frandom.normal(key1, (100, 2))
    model = RealNVP(2, 4, 32, rng_)
    opt = Optimizer(model, learning_rate=0.001, momentum=0.9)
    rng_, best_model, state, loss_values = model.train(init_rng, data, opt.optim, opt.optim_state, 5, 100, verbose=False)
    assert loss_values.shape == (5,) and torch.isfinite(loss_values).all()
    s = best_model.sample(frandom.PRNGKey(124098), 10000)
    assert s.shape == (10000, 2) and torch.isfinite(s).all()
    assert abs(float(s.mean())) < 0.2 and 0.7 < float(s.std()) < 1.3


@pytest.mark.parametrize("tag,bs", [("simple", 100), ("batched", 3)])
def test_nf_proposal_with_realnvp(cuda, tag, bs):
    """TakeGroupSteps + NFProposal(RealNVP) against the golden NFProposal run (both branches of sample_flow's key
    schedule, NF_proposal.py:135-172)."""
    from flowmc_b200 import targets as T
    from flowmc_b200.resource.buffers import Buffer
    from flowmc_b200.resource.kernel.NF_proposal import NFProposal
    from flowmc_b200.resource.logPDF import LogPDF
    from flowmc_b200.resource.states import State
    from flowmc_b200.strategy.take_steps import TakeGroupSteps
    d, L, h = [int(v) for v in G["shape"]]
    p = _params(31, d, L, h)
    m = nvp_model_from_params(p)
    n, n_steps = 6, 7
    res = {"p": Buffer("p", (n, n_steps, d), 1), "l": Buffer("l", (n, n_steps), 1), "a": Buffer("a", (n, n_steps), 1),
           "s": State({"p": "p", "l": "l", "a": "a"}, "s"), "k": NFProposal(m, n_NFproposal_batch_size=bs),
           "logpdf": LogPDF(T.iso_gaussian(0.5, None), n_dims=d)}
    strat = TakeGroupSteps("logpdf", "k", "s", ["p", "l", "a"], n_steps)
    x0 = torch.from_numpy(G["nf_x0"]).cuda()
    new_key, res, last = strat(G["nf_key"], res, x0, None)
    assert np.array_equal(new_key, G[f"nf_{tag}_key"])
    acc = res["a"].data.cpu().numpy()
    assert np.array_equal(acc, G[f"nf_{tag}_acc"])
    assert_close(res["p"].data.cpu().numpy(), G[f"nf_{tag}_pos"], "positions", rtol=3e-5)
    assert_close(res["l"].data.cpu().numpy(), G[f"nf_{tag}_lp"], "log-probs", rtol=3e-5)
    assert torch.equal(last, res["p"].data[:, -1])
    # sharded chains draw the same proposals (global chain keys)
    res2 = {**res, "p": Buffer("p", (3, n_steps, d), 1), "l": Buffer("l", (3, n_steps), 1),
            "a": Buffer("a", (3, n_steps), 1)}
    s2 = TakeGroupSteps("logpdf", "k", "s", ["p", "l", "a"], n_steps)
    s2.set_chain_shard(3, n)
    s2(G["nf_key"], res2, x0[3:], None)
    assert torch.equal(res2["p"].data, res["p"].data[3:])
