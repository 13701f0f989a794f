"""The RealNVP restatement (oracle/realnvp.py; reference src/flowMC/resource/model/nf_model/realNVP.py:102-228 and
resource/model/common.py:68-209) against the committed golden fixture, the reference's own invariants
(test/unit/test_nf.py:24-52) and independent evaluations of the same formulas (torch float64, finite differences)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

from oracle import nf, realnvp as onvp, rng, targets as otargets  # noqa: E402

G = np.load(os.path.join(HERE, "golden", "realnvp_d5.npz"))


def _golden_params():
    from make_golden import realnvp_params
    d, L, h = [int(v) for v in G["shape"]]
    p = realnvp_params(31, d, L, h)
    assert np.array_equal(onvp.flatten(p), G["params_flat"])
    return p


def test_golden_is_frozen():
    p = _golden_params()
    x = G["x"]
    y, ld = onvp.forward(p, x)
    xi, ldi = onvp.inverse(p, x)
    for got, name in ((y, "fwd_y"), (ld, "fwd_logdet"), (xi, "inv_x"), (ldi, "inv_logdet"),
                      (onvp.log_prob(p, x), "log_prob"), (onvp.sample(p, G["sample_key"], 16), "sample")):
        np.testing.assert_allclose(got, G[name], rtol=2e-6, atol=2e-6, err_msg=name)
    loss, g = onvp.loss_and_grads(p, x)
    assert abs(loss - float(G["loss"])) < 1e-5
    np.testing.assert_allclose(onvp.flatten(g, p), G["grad_flat"], rtol=1e-5, atol=1e-6)
    packed = otargets.IsoGaussian.pack(p.n_features, 0.5)
    for tag, bs in (("simple", 100), ("batched", 3)):
        nk, pos, lp, acc, last, dbg = nf.take_group_steps(G["nf_key"], G["nf_x0"], p, "iso_gaussian", packed, 7, bs)
        assert np.array_equal(nk, G[f"nf_{tag}_key"]) and np.array_equal(acc, G[f"nf_{tag}_acc"])
        np.testing.assert_allclose(pos, G[f"nf_{tag}_pos"], rtol=2e-6, atol=2e-6)


def test_reference_invariants_test_nf():
    """test/unit/test_nf.py:24-52: y = model(x), x' = model.inverse(y): x' == x, log_det == -log_det_inv; shapes of
    sample / log_prob.  Holds exactly for the 0 / 1 masks of a freshly built model."""
    key = rng.split(rng.PRNGKey(0), 2)[0]
    p = onvp.init_params(key, 3, 2, 4)
    x = np.array([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]], np.float32)
    y, ld = onvp.forward(p, x)
    xi, ldi = onvp.inverse(p, y)
    assert y.shape == x.shape and ld.shape == (2,)
    np.testing.assert_allclose(xi, x, rtol=1e-5, atol=1e-6)       # jnp.allclose defaults
    np.testing.assert_allclose(ld, -ldi, rtol=1e-5, atol=1e-8)
    s = onvp.sample(p, rng.PRNGKey(0), 2)
    assert s.shape == (2, 3) and onvp.log_prob(p, s).shape == (2,)
    # a model with large weights still inverts (the bijection does not rely on near-identity initialisation)
    p = onvp.init_params(rng.PRNGKey(5), 6, 5, 16)
    p.W1 *= np.float32(40.0)
    x = rng.normal(rng.PRNGKey(6), (50, 6))
    y, ld = onvp.forward(p, x)
    xi, ldi = onvp.inverse(p, y)
    assert np.abs(y - x).max() > 0.1
    np.testing.assert_allclose(xi, x, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(ld, -ldi, rtol=1e-4, atol=1e-6)


def test_masks_and_initialisation():
    """realNVP.py:160-163: ones with the first int(d/2) entries zeroed, flipped on even layers; MLP.__init__
    (common.py:93-107): first-layer weights ~ N(0, 1e-4 / d), everything else uniform(+-1/sqrt(fan_in))."""
    d, L, h = 5, 3, 64
    p = onvp.init_params(rng.PRNGKey(1), d, L, h)
    assert np.array_equal(p.mask[0], [1, 1, 0, 0, 0]) and np.array_equal(p.mask[1], [0, 0, 1, 1, 1])
    assert np.array_equal(p.mask[2], p.mask[0])
    assert abs(p.W1.std() - np.sqrt(1e-4 / d)) < 0.1 * np.sqrt(1e-4 / d)
    assert np.abs(p.b1).max() <= 1 / np.sqrt(d) and np.abs(p.b1).max() > 0.9 / np.sqrt(d)
    assert np.abs(p.W2).max() <= 1 / np.sqrt(h) and np.abs(p.W2).max() > 0.95 / np.sqrt(h)
    assert not np.array_equal(p.W1[0], p.W1[1])                       # scale and shift MLPs use different sub-keys
    q = onvp.init_params(rng.PRNGKey(1), d, L, h)
    assert np.array_equal(onvp.flatten(p), onvp.flatten(q))
    np.testing.assert_array_equal(onvp.flatten(onvp.unflatten(p, onvp.flatten(p))), onvp.flatten(p))
    np.testing.assert_array_equal(onvp.flatten(onvp.init_params(rng.PRNGKey(0), 3, 2, 4)), G["init_key0_flat"])


def test_gradients_against_finite_differences():
    p = _golden_params()
    x = G["x"]
    loss, g = onvp.loss_and_grads(p, x)
    flat, gflat = onvp.flatten(p).astype(np.float64), onvp.flatten(g, p)

    def f(v):
        q = onvp.unflatten(p, v.astype(np.float32))
        # float64 evaluation through the torch path: reuse loss_and_grads' forward
        return onvp.loss_and_grads(q, x)[0]
    r = np.random.default_rng(0)
    nz = np.nonzero(gflat)[0]
    for i in r.choice(nz, 12, replace=False):
        e = np.zeros_like(flat)
        hstep = 1e-3 * max(1.0, abs(flat[i]))
        e[i] = hstep
        fd = (f(flat + e) - f(flat - e)) / (2 * hstep)
        assert abs(fd - gflat[i]) <= 2e-3 * max(abs(gflat[i]), 1e-2), (i, fd, gflat[i])
    # masks, whitening constants and the base distribution get no gradient (stop_gradient)
    q = onvp.unflatten(p, gflat)
    assert not q.mask.any() and not q.data_mean.any() and not q.data_cov.any() and not q.base_cov.any()


def test_train_decays_masks_and_reduces_loss():
    """optax.adamw's weight decay reaches the float masks (zero gradient): after t steps the ones
    are (1 - lr * 1e-4)^t; the zeros stay zero."""
    p = onvp.init_params(rng.PRNGKey(3), 2, 4, 32)
    z = rng.normal(rng.PRNGKey(4), (200, 2))
    data = np.stack([z[:, 0], z[:, 0] ** 2 + np.float32(0.3) * z[:, 1]], axis=1).astype(np.float32)   # a banana
    st = nf.AdamWState(onvp.flatten(p).size)
    key, best, best_st, losses = onvp.train(p, rng.PRNGKey(9), data, st, 1e-2, 15, 100)
    assert losses.min() < losses[0] - 0.05 and best_st.count % 2 == 0 and 0 < best_st.count <= 30
    want = np.float32(1.0)
    for _ in range(best_st.count):
        want = np.float32(want + np.float32(-1e-2) * (np.float32(1e-4) * want))
    ones = best.mask[p.mask == 1]
    np.testing.assert_allclose(ones, want, rtol=1e-6)
    assert (best.mask[p.mask == 0] == 0).all()
