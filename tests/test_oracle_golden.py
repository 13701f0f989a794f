"""CPU tests (no GPU): the oracle against the committed golden vectors, the known-answer tests that
pin it to upstream semantics, and the reference's own invariant tests re-expressed on the oracle
(test/unit/test_nf.py:24-43 forward/inverse consistency; test/unit/test_kernels.py acceptance -> 1
for tiny steps; docs/tutorials/dualmoon.ipynb:65 dual-moon value)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
GOLD = os.path.join(HERE, "golden")

from flowutil import random_params  # noqa: E402
from oracle import flow as oflow, local as olocal, nf, rng, targets as otargets  # noqa: E402


def _load(name):
    return np.load(os.path.join(GOLD, name))


def test_rng_golden():
    g = _load("rng_key42.npz")
    key = rng.PRNGKey(42)
    assert np.array_equal(rng.split(key, 4), g["split"])
    assert np.array_equal(rng.split(key, 2), np.array([[1832780943, 270669613], [64467757, 2916123636]], np.uint32))
    assert np.array_equal(rng.random_bits(key, (16,)), g["bits"])
    assert np.array_equal(rng.uniform(key, (16,)), g["uniform"])
    assert np.array_equal(rng.normal(key, (16,)), g["normal"])
    assert np.array_equal(rng.permutation(key, 2000), g["permutation"])
    assert np.array_equal(rng.choice_with_replacement(key, 1000, 64), g["choice"])
    assert sorted(g["permutation"].tolist()) == list(range(2000))


def test_dual_moon_known_answer():
    """docs/tutorials/dualmoon.ipynb:65: target_dual_moon(zeros(5)) = -218.14496."""
    v = otargets.TARGETS["dual_moon"].logp_grad(np.zeros((1, 5), np.float32), otargets.DualMoon.pack(5, None))[0]
    assert abs(float(v[0]) - (-218.14496)) < 2e-4


def test_flow_golden():
    g = _load("flow_d5.npz")
    d, L, K = (int(v) for v in g["shape"][:3])
    hidden = [int(v) for v in g["shape"][3:]]
    p = random_params(11, d, L, hidden, K)
    assert np.array_equal(nf.flatten(p), g["params_flat"])
    x = g["x"]
    y, ld = oflow.forward(p, x)
    xi, ldi = oflow.inverse(p, x)
    for got, name in ((y, "fwd_y"), (ld, "fwd_logdet"), (xi, "inv_x"), (ldi, "inv_logdet"),
                      (oflow.log_prob(p, x), "log_prob"), (oflow.sample(p, g["sample_key"], 16), "sample")):
        np.testing.assert_allclose(got, g[name], rtol=2e-6, atol=2e-6, err_msg=name)   # BLAS summation order only
    loss, gr = nf.loss_and_grads(p, x)
    assert abs(loss - float(g["loss"])) < 1e-4
    np.testing.assert_allclose(nf.flatten(gr, p), g["grad_flat"], rtol=1e-5, atol=1e-7)


def test_flow_init_golden():
    g = _load("flow_init_key42.npz")
    p = oflow.init_params(rng.PRNGKey(42), 5, 4, [32, 32], 8)
    assert np.array_equal(nf.flatten(p), g["params_flat"])
    # equinox Linear init bounds and the 1e-2/in weight scale (common.py:93-107)
    assert np.abs(p.W[2]).max() <= 1 / np.sqrt(32) and np.abs(p.b[0]).max() <= 1 / np.sqrt(5)
    assert abs(p.W[1].std() - np.sqrt(1e-2 / 32)) < 0.1 * np.sqrt(1e-2 / 32)


def test_flow_invariants_like_reference():
    """test/unit/test_nf.py:38-43: inverse(forward(x)) == x and log-dets cancel (exact inverse while the
    ScalarAffine parameters are at their zero initialisation)."""
    p = oflow.init_params(rng.PRNGKey(1), 6, 5, [32, 32], 8)
    x = (3 * np.random.default_rng(0).standard_normal((200, 6))).astype(np.float32)
    y, ld = oflow.forward(p, x)
    xb, ldb = oflow.inverse(p, y)
    np.testing.assert_allclose(xb, x, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(ld, -ldb, rtol=1e-4, atol=1e-4)
    # log_prob integrates to ~1 in 1-D (importance estimate under the flow's own samples is exactly 1, so use a grid)
    p1 = random_params(3, 1, 2, [8], 8, gain=2.0, affine=0.1, whiten=False)
    grid = np.linspace(-14, 14, 20001, dtype=np.float32)[:, None]
    mass = np.trapezoid(np.exp(oflow.log_prob(p1, grid).astype(np.float64)), grid[:, 0].astype(np.float64))
    assert abs(mass - 1.0) < 2e-3
    # once ScalarAffine trains, `inverse` is no longer the inverse of `forward` (replicated quirk B.6)
    p.scale[:] = 0.3
    p.shift[:] = 0.2
    y, _ = oflow.forward(p, x)
    xb, _ = oflow.inverse(p, y)
    assert np.abs(xb - x).max() > 1e-2


def test_float64_mode_agrees_with_float32():
    p = random_params(11, 5, 3, [16, 16], 8)
    x = (2 * np.random.default_rng(1).standard_normal((50, 5))).astype(np.float32)
    a = oflow.log_prob(p, x)
    with oflow.precision(np.float64):
        b = oflow.log_prob(p, x)
    assert b.dtype == np.float64 and a.dtype == np.float32
    np.testing.assert_allclose(a, b, rtol=1e-4, atol=1e-4)
    lp, g = nf.loss_and_grads(p, x)
    assert abs(lp + b.mean()) < 1e-6 * max(1, abs(lp))      # torch float64 restatement == numpy float64 restatement


def test_local_golden():
    g = _load("local_dualmoon_d5.npz")
    packed = otargets.DualMoon.pack(5, None)
    for kind, kw in (("MALA", dict(step_size=0.1)), ("GRW", dict(step_size=0.3)),
                     ("HMC", dict(step_size=0.05, n_leapfrog=4, condition_matrix=np.eye(5, dtype=np.float32)))):
        k = olocal.make_kernel(kind, **kw)
        nk, pos, lp, acc, last = olocal.take_serial_steps(g["key"], g["x0"], "dual_moon", packed, k, 12)
        assert np.array_equal(nk, g[f"{kind}_key"])
        assert np.array_equal(acc, g[f"{kind}_acc"])
        np.testing.assert_allclose(pos, g[f"{kind}_pos"], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(lp, g[f"{kind}_lp"], rtol=1e-6, atol=1e-5)


def test_nf_global_golden_and_branches():
    g = _load("nf_global_iso_d5.npz")
    p = random_params(21, 5, 2, [16, 16], 8, gain=1.0, affine=0.0, whiten=False)
    assert np.array_equal(nf.flatten(p), g["params_flat"])
    packed = otargets.IsoGaussian.pack(5, 0.5)
    for tag, bs in (("simple", 100), ("batched", 3)):
        nk, pos, lp, acc, last, dbg = nf.take_group_steps(g["key"], g["x0"], p, "iso_gaussian", packed, 7, bs)
        assert np.array_equal(nk, g[f"{tag}_key"])
        assert np.array_equal(acc, g[f"{tag}_acc"])
        np.testing.assert_allclose(pos, g[f"{tag}_pos"], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(dbg["lp_nf_prop"], g[f"{tag}_lp_nf"], rtol=1e-5, atol=1e-5)
    # the two branches of sample_flow use different key schedules (NF_proposal.py:135-172)
    assert not np.allclose(g["simple_proposals"], g["batched_proposals"])
    # accepted proposals are copied verbatim; rejected steps repeat the previous position
    pos, acc, prop = g["simple_pos"], g["simple_acc"], g["simple_proposals"]
    for c in range(pos.shape[0]):
        prev = g["x0"][c]
        for t in range(pos.shape[1]):
            want = prop[c, t] if acc[c, t] else prev
            assert np.array_equal(pos[c, t], want)
            prev = pos[c, t]


def test_optimizer_restatement_properties():
    """optax.chain(clip_by_global_norm(1), adamw): the first step moves every weight by ~lr against the
    gradient sign; a gradient below the clip norm is left unscaled; zero-gradient leaves only decay."""
    n = 64
    st = nf.AdamWState(n)
    p0 = np.ones(n, np.float32)
    g = np.zeros(n, np.float32)
    g[:32] = 10.0
    p1, gnorm = nf.clip_adamw(p0, g, st, 1e-2)
    assert abs(gnorm - np.sqrt(32 * 100.0)) < 1e-3
    np.testing.assert_allclose(p1[:32], 1.0 - 1e-2 * (1.0 + 1e-4), rtol=1e-5)     # m_hat/sqrt(v_hat) = 1 on step 1
    np.testing.assert_allclose(p1[32:], 1.0 - 1e-2 * 1e-4, rtol=1e-6)              # weight decay only
    st2 = nf.AdamWState(n)
    small = np.full(n, 1e-3, np.float32)
    nf.clip_adamw(p0, small, st2, 1e-2)
    np.testing.assert_allclose(st2.mu, 0.1 * small, rtol=1e-6)                      # not clipped


def test_select_training_data_layout():
    """train_model.py:66-81: finite rows, last `history_window` per chain, chain-major flattening."""
    n_chains, n_total, d = 4, 9, 2
    buf = np.full((n_chains, n_total, d), -np.inf, np.float32)
    buf[:, :6] = np.arange(n_chains * 6 * d, dtype=np.float32).reshape(n_chains, 6, d)
    key, tkey, data, idx = nf.select_training_data(rng.PRNGKey(0), buf, 50, 3)
    pop = buf[:, 3:6].reshape(-1, d)
    assert np.array_equal(data, pop[idx]) and idx.max() < 12 and data.shape == (50, d)


def test_adam_optimization_restatement_meets_the_reference_invariants():
    """test/unit/test_strategies.py:40-78 on the oracle's AdamOptimization (oracle/optimization.py)."""
    from oracle import optimization as oopt, rng, targets as O
    key = rng.PRNGKey(42)
    key, sub = rng.split(key)
    x0 = (rng.normal(sub, (20, 2)) * 1 + 10).astype(np.float32)
    data = O.IsoGaussian.pack(2, 0.5, np.arange(2))
    new_key, x, lp = oopt.adam_optimize(key, "iso_gaussian", data, x0, n_steps=100, learning_rate=5e-2, noise_level=0.0)
    assert x.shape == (20, 2) and lp.shape == (20,)
    assert np.all(x.mean(axis=1) < x0.mean(axis=1)) and np.all(np.isfinite(lp))
    assert np.array_equal(new_key, rng.split(key)[0])
    # Adam moves every coordinate by at most ~lr per step towards the optimum: 100 steps of 0.05
    assert np.all((x0 - x) > 3.5) and np.all((x0 - x) < 5.01)
    # the box projection holds at every step, noise or not; same key -> same result
    _, xb, _ = oopt.adam_optimize(key, "iso_gaussian", data, x0, n_steps=30, noise_level=10.0, bounds=[[9.0, 10.5]])
    _, xb2, _ = oopt.adam_optimize(key, "iso_gaussian", data, x0, n_steps=30, noise_level=10.0, bounds=[[9.0, 10.5]])
    assert xb.min() >= 9.0 and xb.max() <= 10.5 and np.array_equal(xb, xb2)


def test_parallel_tempering_restatement_invariants():
    """oracle/parallel_tempering.py: shapes and invariants of test/unit/test_strategies.py:337-520."""
    from oracle import parallel_tempering as opt, rng, targets as O
    key = rng.PRNGKey(42)
    key, sub = rng.split(key)
    x0 = rng.normal(sub, (7, 3))
    key, sub = rng.split(key)
    tp = rng.normal(sub, (7, 4, 3))
    temps = (np.arange(5) + 1.0).astype(np.float32)
    data = O.IsoGaussian.pack(3, 0.5, np.arange(3))
    k1, p0, tpos, t2, accs = opt.parallel_tempering(key, x0, tp, temps, "iso_gaussian", data, 4, 1.0)
    k2, p0b, _, _, _ = opt.parallel_tempering(key, x0, tp, temps, "iso_gaussian", data, 4, 1.0)
    assert p0.shape == (7, 3) and tpos.shape == (7, 4, 3) and accs.shape == (7, 4)
    assert np.array_equal(p0, p0b) and np.array_equal(k1, k2)                 # deterministic in the key
    assert t2[0] == temps[0] and t2[-1] == temps[-1]                          # the ladder's ends never move
    # equal acceptance on every rung leaves the ladder unchanged (test_adapt_temperatures)
    t = (np.arange(5) * 0.3 + 1).astype(np.float32)
    assert np.allclose(opt.adapt_temperature(t, np.ones((7, 4), np.float32)), t)
    # an exchange only permutes the rungs of a chain
    pos = np.concatenate([x0[:, None, :], tp], axis=1)
    out, lp, acc, _, _ = opt.exchange(sub, pos, "iso_gaussian", data, t)
    assert np.allclose(np.sort(out.sum(axis=2), axis=1), np.sort(pos.sum(axis=2), axis=1))
    # not training: buffers untouched
    _, _, tpos_nt, t_nt, _ = opt.parallel_tempering(key, x0, tp, temps, "iso_gaussian", data, 4, 1.0, training=False)
    assert np.array_equal(tpos_nt, tp) and np.array_equal(t_nt, temps)


def test_strategies_golden():
    """oracle AdamOptimization / ParallelTempering against the committed vectors (tests/golden/strategies_iso.npz)."""
    from oracle import optimization as oopt, parallel_tempering as opt_pt, targets as O
    from oracle import rng as orng
    g = _load("strategies_iso.npz")
    key = orng.split(orng.PRNGKey(42))[0]
    d2 = O.IsoGaussian.pack(2, 0.5, np.arange(2))
    k, x, lp = oopt.adam_optimize(key, "iso_gaussian", d2, g["adam_x0"], n_steps=100, learning_rate=5e-2, noise_level=0.0)
    assert np.array_equal(k, g["adam_key"])
    np.testing.assert_allclose(x, g["adam_x"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(lp, g["adam_lp"], rtol=1e-6, atol=1e-5)
    k, x, lp = oopt.adam_optimize(key, "iso_gaussian", d2, g["adam_x0"], n_steps=30, learning_rate=1e-2, noise_level=10.0,
                                  bounds=[[9.0, 10.5]])
    np.testing.assert_allclose(x, g["adam_noisy_x"], rtol=1e-6, atol=1e-6)
    d3 = O.IsoGaussian.pack(3, 0.5, np.arange(3))
    temps = (np.arange(5) + 1.0).astype(np.float32)
    k, p0, tp, t, acc = opt_pt.parallel_tempering(g["pt_key_in"], g["pt_x0"], g["pt_tempered_in"], temps, "iso_gaussian",
                                                  d3, 4, 1.0)
    assert np.array_equal(k, g["pt_key"]) and np.array_equal(acc, g["pt_accepts"])
    np.testing.assert_allclose(p0, g["pt_positions"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(tp, g["pt_tempered"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(t, g["pt_temperatures"], rtol=1e-6)
