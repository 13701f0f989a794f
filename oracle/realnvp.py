"""Oracle: RealNVP (forward / inverse / log_prob / sample / init / loss gradients).  TEST ONLY.

Restates, vectorised over samples in numpy fp32 (all paths under src/flowMC/resource/model/):
  * RealNVP.__init__ / forward / inverse / sample / log_prob       nf_model/realNVP.py:102-228
  * MLPAffine.forward / inverse                                    common.py:171-209
  * MaskedCouplingLayer.forward / inverse                          common.py:150-168
  * MLP.__init__ / __call__ with the DEFAULT activation jax.nn.relu (RealNVP builds MLP([d, h, d], key) without an
    activation argument, realNVP.py:164-165)                        common.py:68-124
  * Gaussian.sample (base distribution)                            common.py:285-293
equinox.nn.Linear's default init (uniform(-1/sqrt(in), 1/sqrt(in)) for weight and bias, wkey, bkey = split(key, 2)) is
restated from equinox 0.11.11.

Reference quirks kept on purpose (they change trained models, so parity needs them):
  * the coupling masks are FLOAT arrays (jnp.ones / .at[].set(0), realNVP.py:160-163) held as model leaves behind
    stop_gradient (common.py:142-144): they get zero gradients, and optax.adamw's weight decay -- applied to every
    leaf that has a gradient, SURVEY.md B.4 -- shrinks the ones towards zero, a factor (1 - lr * 1e-4) per step.  The
    layer formulas are therefore evaluated with a general float mask m: cond = x * m, y = (1 - m) * b(x) + m * x,
    log_det = sum((1 - m) * scale);
  * log_prob adds multivariate_normal.logpdf(y, zeros, eye) -- literal zeros / eye, NOT base_dist (realNVP.py:218-220),
    while sample draws from base_dist, whose covariance weight decay also shrinks;
  * make_layer's split(key, 3) keeps scale_subkey = [1], shift_subkey = [2] (realNVP.py:159).

Parameter container ``NVPParams`` mirrors the flat device blob (include/flowmc_b200.h FlowmcRealNVPDesc).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import rng

F32 = np.float32
_LOG_2PI = F32(np.log(2 * np.pi))


@dataclass
class NVPParams:
    n_features: int
    n_layers: int
    n_hidden: int
    # per layer (leading axis L); weights are (out, in) like equinox.nn.Linear; index 0 = scale MLP, 1 = shift MLP
    W1: np.ndarray = None      # [2, L, h, d]
    b1: np.ndarray = None      # [2, L, h]
    W2: np.ndarray = None      # [2, L, d, h]
    b2: np.ndarray = None      # [2, L, d]
    mask: np.ndarray = None    # [L, d] float: 1 = conditioning (unchanged), 0 = transformed
    data_mean: np.ndarray = None
    data_cov: np.ndarray = None
    base_mean: np.ndarray = None
    base_cov: np.ndarray = None

    def copy(self) -> "NVPParams":
        return NVPParams(self.n_features, self.n_layers, self.n_hidden, self.W1.copy(), self.b1.copy(), self.W2.copy(),
                         self.b2.copy(), self.mask.copy(), self.data_mean.copy(), self.data_cov.copy(),
                         self.base_mean.copy(), self.base_cov.copy())


def make_mask(n_features: int, layer: int) -> np.ndarray:
    """realNVP.py:160-163: ones with the first int(d/2) entries zero; flipped on even layers."""
    m = np.ones(n_features, F32)
    m[: int(n_features / 2)] = 0
    return (F32(1) - m) if layer % 2 == 0 else m


def _mlp_init(key, shape, scale=1e-4):
    """MLP.__init__ (common.py:83-107) -> ([W...], [b...])."""
    Ws, bs = [], []
    k = key
    for i in range(len(shape) - 2):
        k3 = rng.split(k, 3)
        k, sub1, sub2 = k3[0], k3[1], k3[2]
        _, bkey = rng.split(sub1, 2)
        lim = F32(1.0 / np.sqrt(shape[i]))
        bs.append(rng.uniform(bkey, (shape[i + 1],), -lim, lim))
        w = rng.normal(sub2, (shape[i + 1], shape[i]))
        Ws.append((w * np.sqrt(F32(scale / shape[i]))).astype(F32))
    sub = rng.split(k, 2)[1]
    wkey, bkey = rng.split(sub, 2)
    lim = F32(1.0 / np.sqrt(shape[-2]))
    Ws.append(rng.uniform(wkey, (shape[-1], shape[-2]), -lim, lim))
    bs.append(rng.uniform(bkey, (shape[-1],), -lim, lim))
    return Ws, bs


def init_params(key, n_features, n_layers, n_hidden) -> NVPParams:
    """RealNVP.__init__ (realNVP.py:131-170): keys = split(key, n_layers); per layer split(key, 3)[1:] seed the
    scale and shift MLPs."""
    d, h, L = n_features, n_hidden, n_layers
    p = NVPParams(d, L, h, np.zeros((2, L, h, d), F32), np.zeros((2, L, h), F32), np.zeros((2, L, d, h), F32),
                  np.zeros((2, L, d), F32), np.stack([make_mask(d, l) for l in range(L)]), np.zeros(d, F32),
                  np.eye(d, dtype=F32), np.zeros(d, F32), np.eye(d, dtype=F32))
    keys = rng.split(key, L)
    for l in range(L):
        k3 = rng.split(keys[l], 3)
        for which, sub in ((0, k3[1]), (1, k3[2])):
            Ws, bs = _mlp_init(sub, [d, h, d])
            p.W1[which, l], p.b1[which, l], p.W2[which, l], p.b2[which, l] = Ws[0], bs[0], Ws[1], bs[1]
    return p


def _mlp(p: NVPParams, which: int, layer: int, u: np.ndarray) -> np.ndarray:
    """common.py:109-112 with relu: u [n, d] -> [n, d]."""
    hdn = (u @ p.W1[which, layer].T + p.b1[which, layer]).astype(F32)
    hdn = np.maximum(hdn, F32(0))
    return (hdn @ p.W2[which, layer].T + p.b2[which, layer]).astype(F32)


def _coupling(p: NVPParams, layer: int, x: np.ndarray, inverse: bool):
    """MaskedCouplingLayer(MLPAffine) forward / inverse (common.py:150-168,186-209)."""
    m = p.mask[layer].astype(F32)
    cond = (x * m).astype(F32)
    scale = np.tanh(_mlp(p, 0, layer, cond)).astype(F32)          # * dt, dt = 1
    shift = _mlp(p, 1, layer, cond)
    if not inverse:
        y = ((x + shift) * np.exp(scale)).astype(F32)
        ld = scale
    else:
        y = (x * np.exp(-scale) - shift).astype(F32)
        ld = -scale
    y = ((F32(1) - m) * y + m * x).astype(F32)
    return y, np.sum((F32(1) - m) * ld, axis=-1, dtype=F32).astype(F32)


def forward(p: NVPParams, x: np.ndarray):
    x = np.asarray(x, F32)
    ld = np.zeros(x.shape[0], F32)
    for l in range(p.n_layers):
        x, li = _coupling(p, l, x, False)
        ld = (ld + li).astype(F32)
    return x, ld


def inverse(p: NVPParams, x: np.ndarray):
    x = np.asarray(x, F32)
    ld = np.zeros(x.shape[0], F32)
    for l in reversed(range(p.n_layers)):
        x, li = _coupling(p, l, x, True)
        ld = (ld + li).astype(F32)
    return x, ld


def log_prob(p: NVPParams, x: np.ndarray):
    """realNVP.py:214-221."""
    x = ((np.asarray(x, F32) - p.data_mean) / np.sqrt(np.diag(p.data_cov))).astype(F32)
    y, ld = forward(p, x)
    base = (F32(-0.5) * np.sum(y * y, axis=-1, dtype=F32) - F32(p.n_features / 2) * _LOG_2PI).astype(F32)
    return (ld + base).astype(F32)


def sample_from_z(p: NVPParams, z: np.ndarray):
    """realNVP.py:208-212 from standard normals z: base sample (diagonal base covariance) -> inverse -> un-whiten."""
    s = (p.base_mean + z * np.sqrt(np.diag(p.base_cov))).astype(F32)
    x, _ = inverse(p, s)
    return (x * np.sqrt(np.diag(p.data_cov)) + p.data_mean).astype(F32)


def sample(p: NVPParams, key, n: int):
    return sample_from_z(p, rng.normal(key, (n, p.n_features)))


# ------------------------------------------------------------------------------ float64 autograd of the same formulas
def loss_and_grads(p: NVPParams, x: np.ndarray):
    """NFModel.loss_fn (nf_model/base.py:98-100) = -mean(log_prob) and its gradient w.r.t. the MLP weights (torch
    float64 autograd of the restated formulas; masks / data_mean / data_cov sit behind stop_gradient)."""
    import torch
    dt = torch.float64
    W1 = torch.tensor(p.W1, dtype=dt, requires_grad=True)
    b1 = torch.tensor(p.b1, dtype=dt, requires_grad=True)
    W2 = torch.tensor(p.W2, dtype=dt, requires_grad=True)
    b2 = torch.tensor(p.b2, dtype=dt, requires_grad=True)
    mask = torch.tensor(p.mask, dtype=dt)
    xx = (torch.tensor(np.asarray(x, F32), dtype=dt) - torch.tensor(p.data_mean, dtype=dt)) / torch.sqrt(
        torch.tensor(np.diag(p.data_cov).copy(), dtype=dt))
    ld = torch.zeros(xx.shape[0], dtype=dt)
    for l in range(p.n_layers):
        m = mask[l]
        cond = xx * m
        outs = []
        for w in range(2):
            hdn = torch.relu(cond @ W1[w, l].T + b1[w, l])
            outs.append(hdn @ W2[w, l].T + b2[w, l])
        scale = torch.tanh(outs[0])
        y = (xx + outs[1]) * torch.exp(scale)
        xx = (1 - m) * y + m * xx
        ld = ld + ((1 - m) * scale).sum(-1)
    d = p.n_features
    lp = ld - 0.5 * (xx * xx).sum(-1) - d / 2 * float(np.log(2 * np.pi))
    loss = -lp.mean()
    g = torch.autograd.grad(loss, [W1, b1, W2, b2])
    return float(loss.detach()), dict(W1=g[0].numpy().astype(F32), b1=g[1].numpy().astype(F32),
                                      W2=g[2].numpy().astype(F32), b2=g[3].numpy().astype(F32))


def flatten(p_or_g, p: NVPParams = None) -> np.ndarray:
    """Flat float32 vector in the device blob order (FlowmcRealNVPDesc, without its alignment padding): per layer
    [W1s, b1s, W2s, b2s, W1t, b1t, W2t, b2t, mask], then data_mean, data_cov, base_mean, base_cov.  For a grads dict
    the masks and the tail are zero."""
    if isinstance(p_or_g, dict):
        g, ref = p_or_g, p
        mask = np.zeros_like(ref.mask)
        tail = [np.zeros_like(ref.data_mean), np.zeros_like(ref.data_cov), np.zeros_like(ref.base_mean),
                np.zeros_like(ref.base_cov)]
        W1, b1, W2, b2 = g["W1"], g["b1"], g["W2"], g["b2"]
    else:
        ref = p_or_g
        mask = ref.mask
        tail = [ref.data_mean, ref.data_cov, ref.base_mean, ref.base_cov]
        W1, b1, W2, b2 = ref.W1, ref.b1, ref.W2, ref.b2
    parts = []
    for l in range(ref.n_layers):
        for w in range(2):
            parts += [W1[w, l].reshape(-1), b1[w, l], W2[w, l].reshape(-1), b2[w, l]]
        parts.append(mask[l])
    parts += [t.reshape(-1) for t in tail]
    return np.concatenate(parts).astype(F32)


def unflatten(p: NVPParams, flat: np.ndarray) -> NVPParams:
    q = p.copy()
    d, h = p.n_features, p.n_hidden
    o = 0
    for l in range(p.n_layers):
        for w in range(2):
            q.W1[w, l] = flat[o:o + h * d].reshape(h, d); o += h * d
            q.b1[w, l] = flat[o:o + h]; o += h
            q.W2[w, l] = flat[o:o + d * h].reshape(d, h); o += d * h
            q.b2[w, l] = flat[o:o + d]; o += d
        q.mask[l] = flat[o:o + d]; o += d
    q.data_mean = flat[o:o + d].copy(); o += d
    q.data_cov = flat[o:o + d * d].reshape(d, d).copy(); o += d * d
    q.base_mean = flat[o:o + d].copy(); o += d
    q.base_cov = flat[o:o + d * d].reshape(d, d).copy(); o += d * d
    assert o == flat.size
    return q


def train(p: NVPParams, rng_key, data, st, lr, num_epochs, batch_size, momentum=0.9):
    """NFModel.train (nf_model/base.py:153-210) with the Optimizer's chain (oracle.nf.clip_adamw); same contract as
    oracle.nf.train."""
    from . import nf
    st = st.copy()
    data = np.asarray(data, F32)
    N = data.shape[0]
    q = p.copy()
    q.data_mean = data.mean(axis=0, dtype=F32).astype(F32)
    q.data_cov = np.atleast_2d(np.cov(data.T.astype(np.float64))).astype(F32)
    best, best_st, best_loss = p, st.copy(), 1e9
    losses = np.zeros(num_epochs, F32)
    key = np.asarray(rng_key, np.uint32)
    for e in range(num_epochs):
        ks = rng.split(key, 2)
        key, in_key = ks[0], ks[1]
        steps = N // batch_size
        value = 1e9
        if steps > 0:
            perm = rng.permutation(in_key, N)[:steps * batch_size].reshape(steps, batch_size)
            batches = [data[idx] for idx in perm]
        else:
            batches = [data]
        for xb in batches:
            value, g = loss_and_grads(q, xb)
            flat, _ = nf.clip_adamw(flatten(q), flatten(g, q), st, lr, b1=momentum)
            q = unflatten(q, flat)
        losses[e] = value
        if losses[e] < best_loss:
            best, best_st, best_loss = q.copy(), st.copy(), losses[e]
    return key, best, best_st, losses
