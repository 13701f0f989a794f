"""Oracle: MaskedCouplingRQSpline (forward / inverse / log_prob / sample / init) in numpy fp32.  TEST ONLY.

Restates, vectorised over samples:
  * _normalize_bin_sizes / _normalize_knot_slopes      rqSpline.py:20-39
  * _rational_quadratic_spline_fwd / _inv, _safe_quadratic_root   rqSpline.py:42-239
  * RQSpline.get_params                                 rqSpline.py:310-338
  * MLP.__init__/__call__ (tanh conditioner)            common.py:68-124, rqSpline.py:428-433
  * MaskedCouplingLayer.forward/inverse                 common.py:150-168
  * ScalarAffine                                        common.py:211-240
  * Gaussian.log_prob/sample (cov = I)                  common.py:285-293
  * MaskedCouplingRQSpline.__init__/forward/inverse/sample/log_prob   rqSpline.py:392-504
(all paths under src/flowMC/resource/model/).  equinox.nn.Linear's default init
(uniform(-1/sqrt(in), 1/sqrt(in)) for weight and bias, keys = split(key, 2)) is restated from
equinox 0.11.11; jax.nn.softmax/softplus from jax 0.5.0.

Parameter container: ``FlowParams`` -- per-layer arrays stacked on a leading layer axis, the same
quantities the device blob holds (include/flowmc_b200.h FlowmcFlowDesc documents the flat layout).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import rng

import contextlib

F32 = np.float32
MIN_BIN = F32(1e-4)
MIN_SLOPE = F32(1e-4)
_LOG_2PI = F32(np.log(2 * np.pi))


@contextlib.contextmanager
def precision(dtype):
    """Evaluate the same formulas in another dtype (float64): every cast in this module goes through
    the module-level ``F32``.  Tests use |f32 result - f64 result| as the fp32 noise floor of an
    expression (the inverse spline root is ill-conditioned near small discriminants)."""
    global F32, MIN_BIN, MIN_SLOPE, _LOG_2PI
    old = (F32, MIN_BIN, MIN_SLOPE, _LOG_2PI)
    F32 = dtype
    MIN_BIN, MIN_SLOPE, _LOG_2PI = dtype(np.float32(1e-4)), dtype(np.float32(1e-4)), dtype(np.log(2 * np.pi))
    try:
        yield
    finally:
        F32, MIN_BIN, MIN_SLOPE, _LOG_2PI = old


@dataclass
class FlowParams:
    n_features: int
    n_layers: int
    hidden: list
    num_bins: int
    # per layer (leading axis L): weights are (out, in) like equinox.nn.Linear
    W: list = field(default_factory=list)   # list over MLP linear index i of [L, out_i, in_i]
    b: list = field(default_factory=list)   # list over i of [L, out_i]
    scale: np.ndarray = None                # [L]   ScalarAffine.scale
    shift: np.ndarray = None                # [L]
    data_mean: np.ndarray = None            # [d]
    data_cov: np.ndarray = None             # [d, d]
    base_mean: np.ndarray = None            # [d]
    base_cov: np.ndarray = None             # [d, d]
    range_min: float = -10.0
    range_max: float = 10.0

    def mask(self, layer: int) -> np.ndarray:
        """True = conditioning (unchanged) features, rqSpline.py:434."""
        return ((np.arange(self.n_features) + layer) % 2).astype(bool)

    def copy(self) -> "FlowParams":
        return FlowParams(self.n_features, self.n_layers, list(self.hidden), self.num_bins,
                          [w.copy() for w in self.W], [b.copy() for b in self.b], self.scale.copy(),
                          self.shift.copy(), self.data_mean.copy(), self.data_cov.copy(), self.base_mean.copy(),
                          self.base_cov.copy(), self.range_min, self.range_max)


def init_params(key, n_features, n_layers, hidden, num_bins, spline_range=(-10.0, 10.0)) -> FlowParams:
    """MaskedCouplingRQSpline.__init__ (rqSpline.py:392-443) + MLP.__init__ (common.py:83-107)."""
    d = n_features
    shape = [d] + list(hidden) + [d * (3 * num_bins + 1)]
    n_lin = len(shape) - 1
    Ws = [np.zeros((n_layers, shape[i + 1], shape[i]), F32) for i in range(n_lin)]
    bs = [np.zeros((n_layers, shape[i + 1]), F32) for i in range(n_lin)]
    keys = rng.split(key, n_layers)
    for l in range(n_layers):
        k = keys[l]
        for i in range(n_lin - 1):
            k3 = rng.split(k, 3)
            k, sub1, sub2 = k3[0], k3[1], k3[2]
            wkey, bkey = rng.split(sub1, 2)
            lim = F32(1.0 / np.sqrt(shape[i]))
            bs[i][l] = rng.uniform(bkey, (shape[i + 1],), -lim, lim)
            w = rng.normal(sub2, (shape[i + 1], shape[i]))
            Ws[i][l] = (w * np.sqrt(F32(1e-2 / shape[i]))).astype(F32)  # jnp.sqrt(python float) -> f32 sqrt
        k2 = rng.split(k, 2)
        sub = k2[1]
        wkey, bkey = rng.split(sub, 2)
        lim = F32(1.0 / np.sqrt(shape[-2]))
        Ws[-1][l] = rng.uniform(wkey, (shape[-1], shape[-2]), -lim, lim)
        bs[-1][l] = rng.uniform(bkey, (shape[-1],), -lim, lim)
    return FlowParams(d, n_layers, list(hidden), num_bins, Ws, bs, np.zeros(n_layers, F32), np.zeros(n_layers, F32),
                      np.zeros(d, F32), np.eye(d, dtype=F32), np.zeros(d, F32), np.eye(d, dtype=F32),
                      float(spline_range[0]), float(spline_range[1]))


# ------------------------------------------------------------------------------------------- MLP
def mlp(p: FlowParams, layer: int, x: np.ndarray, return_hidden=False):
    """common.py:109-112 with tanh activations: x [n, d] -> [n, d*(3K+1)]."""
    h = x.astype(F32)
    hs = []
    n_lin = len(p.W)
    for i in range(n_lin):
        h = (h @ p.W[i][layer].T + p.b[i][layer]).astype(F32)
        if i < n_lin - 1:
            h = np.tanh(h).astype(F32)
            hs.append(h)
    return (h, hs) if return_hidden else h


def _softmax(u):
    m = u.max(axis=-1, keepdims=True)
    e = np.exp(u - m).astype(F32)
    return (e / e.sum(axis=-1, keepdims=True, dtype=F32)).astype(F32)


def _softplus(u):
    return np.logaddexp(u, F32(0)).astype(F32)


def get_params(p: FlowParams, raw: np.ndarray):
    """rqSpline.py:310-338: raw [n, d*(3K+1)] -> x_pos, y_pos, knot_slopes each [n, d, K+1]."""
    K = p.num_bins
    n = raw.shape[0]
    r = raw.reshape(n, p.n_features, 3 * K + 1)
    uw, uh, us = r[..., :K], r[..., K:2 * K], r[..., 2 * K:]
    rmin, rmax = F32(p.range_min), F32(p.range_max)
    size = F32(rmax - rmin)
    scale = F32(size - F32(K) * MIN_BIN)
    bw = (_softmax(uw) * scale + MIN_BIN).astype(F32)
    bh = (_softmax(uh) * scale + MIN_BIN).astype(F32)
    xp = (rmin + np.cumsum(bw[..., :-1], axis=-1, dtype=F32)).astype(F32)
    yp = (rmin + np.cumsum(bh[..., :-1], axis=-1, dtype=F32)).astype(F32)
    pad_lo = np.full(r.shape[:-1] + (1,), rmin, F32)
    pad_hi = np.full(r.shape[:-1] + (1,), rmax, F32)
    x_pos = np.concatenate([pad_lo, xp, pad_hi], axis=-1)
    y_pos = np.concatenate([pad_lo, yp, pad_hi], axis=-1)
    offset = F32(np.float32(np.log(np.exp(np.float32(1.0) - np.float32(1e-4)) - np.float32(1.0))))
    slopes = (_softplus(us + offset) + MIN_SLOPE).astype(F32)
    return x_pos, y_pos, slopes


def _select_bin(v, pos):
    """one-hot bin of v in pos [..., K+1]; first bin if none (rqSpline.py:63-72)."""
    inb = (v[..., None] >= pos[..., :-1]) & (v[..., None] < pos[..., 1:])
    none = ~inb.any(axis=-1)
    inb[..., 0] |= none
    return inb


def _pick(onehot, arr):
    left = np.sum(onehot * arr[..., :-1], axis=-1, dtype=F32)
    right = np.sum(onehot * arr[..., 1:], axis=-1, dtype=F32)
    return left.astype(F32), right.astype(F32)


def spline_fwd(x, x_pos, y_pos, slopes):
    """rqSpline.py:42-128, elementwise over leading dims."""
    x = x.astype(F32)
    below = x <= x_pos[..., 0]
    above = x >= x_pos[..., -1]
    oh = _select_bin(x, x_pos)
    xl, xr = _pick(oh, x_pos)
    yl, yr = _pick(oh, y_pos)
    dl, dr = _pick(oh, slopes)
    with np.errstate(all="ignore"):
        bw = xr - xl
        bh = yr - yl
        s = (bh / bw).astype(F32)
        z = np.clip((x - xl) / bw, F32(0), F32(1)).astype(F32)
        sq_z = z * z
        z1mz = z - sq_z
        sq_1mz = (F32(1) - z) ** 2
        st = dr + dl - F32(2) * s
        num = bh * (s * sq_z + dl * z1mz)
        den = s + st * z1mz
        y = (yl + num / den).astype(F32)
        logdet = (F32(2) * np.log(s) + np.log(dr * sq_z + F32(2) * s * z1mz + dl * sq_1mz) - F32(2) * np.log(den)).astype(F32)
        y = np.where(below, (x - x_pos[..., 0]) * slopes[..., 0] + y_pos[..., 0], y)
        y = np.where(above, (x - x_pos[..., -1]) * slopes[..., -1] + y_pos[..., -1], y)
        logdet = np.where(below, np.log(slopes[..., 0]), logdet)
        logdet = np.where(above, np.log(slopes[..., -1]), logdet)
    return y.astype(F32), logdet.astype(F32)


def _safe_quadratic_root(a, b, c):
    """rqSpline.py:131-155."""
    with np.errstate(all="ignore"):
        disc = (b * b - F32(4) * a * c).astype(F32)
        sq = np.sqrt(np.maximum(disc, np.finfo(F32).tiny)).astype(F32)
        sq = np.where(disc > 0, sq, F32(0)).astype(F32)
        num = np.where(b >= 0, F32(2) * c, -b + sq)
        den = np.where(b >= 0, -b - sq, F32(2) * a)
        return (num / den).astype(F32)


def spline_inv(y, x_pos, y_pos, slopes):
    """rqSpline.py:158-239."""
    y = y.astype(F32)
    below = y <= y_pos[..., 0]
    above = y >= y_pos[..., -1]
    oh = _select_bin(y, y_pos)
    xl, xr = _pick(oh, x_pos)
    yl, yr = _pick(oh, y_pos)
    dl, dr = _pick(oh, slopes)
    with np.errstate(all="ignore"):
        bw = xr - xl
        bh = yr - yl
        s = (bh / bw).astype(F32)
        w = np.clip((y - yl) / bh, F32(0), F32(1)).astype(F32)
        st = dr + dl - F32(2) * s
        c = -s * w
        b = dl - st * w
        a = s - b
        z = np.clip(_safe_quadratic_root(a, b, c), F32(0), F32(1)).astype(F32)
        x = (bw * z + xl).astype(F32)
        sq_z = z * z
        z1mz = z - sq_z
        sq_1mz = (F32(1) - z) ** 2
        den = s + st * z1mz
        logdet = (-F32(2) * np.log(s) - np.log(dr * sq_z + F32(2) * s * z1mz + dl * sq_1mz) + F32(2) * np.log(den)).astype(F32)
        x = np.where(below, (y - y_pos[..., 0]) / slopes[..., 0] + x_pos[..., 0], x)
        x = np.where(above, (y - y_pos[..., -1]) / slopes[..., -1] + x_pos[..., -1], x)
        logdet = np.where(below, -np.log(slopes[..., 0]), logdet)
        logdet = np.where(above, -np.log(slopes[..., -1]), logdet)
    return x.astype(F32), logdet.astype(F32)


# ---------------------------------------------------------------------------------------- layers
def _coupling(p, layer, x, inverse):
    """MaskedCouplingLayer(RQSpline) forward/inverse (common.py:150-168)."""
    m = p.mask(layer)
    cond = (x * m.astype(F32)).astype(F32)
    raw = mlp(p, layer, cond)
    xp, yp, sl = get_params(p, raw)
    t, ld = (spline_inv if inverse else spline_fwd)(x, xp, yp, sl)
    mf = m.astype(F32)
    y = ((F32(1) - mf) * t + mf * x).astype(F32)
    logdet = np.sum((F32(1) - mf) * ld, axis=-1, dtype=F32).astype(F32)
    return y, logdet


def forward(p: FlowParams, x: np.ndarray):
    """rqSpline.py:450-468: layers 0..L-1, each [ScalarAffine (all-False mask), RQSpline coupling]."""
    x = np.asarray(x, F32)
    d = p.n_features
    logdet = np.zeros(x.shape[0], F32)
    for l in range(p.n_layers):
        x = ((x + p.shift[l]) * np.exp(p.scale[l])).astype(F32)
        logdet = (logdet + np.sum(np.full(d, p.scale[l], F32), dtype=F32)).astype(F32)
        x, ld = _coupling(p, l, x, inverse=False)
        logdet = (logdet + ld).astype(F32)
    return x, logdet


def inverse(p: FlowParams, x: np.ndarray):
    """rqSpline.py:470-488: layers L-1..0, each [ScalarAffine.inverse, RQSpline coupling inverse]
    (the within-layer order is NOT swapped -- replicated as written)."""
    x = np.asarray(x, F32)
    d = p.n_features
    logdet = np.zeros(x.shape[0], F32)
    for l in reversed(range(p.n_layers)):
        x = (x * np.exp(-p.scale[l]) - p.shift[l]).astype(F32)
        logdet = (logdet + np.sum(np.full(d, -p.scale[l], F32), dtype=F32)).astype(F32)
        x, ld = _coupling(p, l, x, inverse=True)
        logdet = (logdet + ld).astype(F32)
    return x, logdet


def base_log_prob(p: FlowParams, y: np.ndarray):
    """Gaussian.log_prob = multivariate_normal.logpdf(y, mean, cov) (common.py:285-286), Cholesky form."""
    d = p.n_features
    L = np.linalg.cholesky(p.base_cov.astype(np.float64)).astype(F32)
    r = np.linalg.solve(L.astype(np.float64), (y - p.base_mean).astype(np.float64).T).T.astype(F32)
    return (F32(-0.5) * np.sum(r * r, axis=-1, dtype=F32) - F32(d / 2) * _LOG_2PI
            - np.sum(np.log(np.diag(L)), dtype=F32)).astype(F32)


def log_prob(p: FlowParams, x: np.ndarray):
    """rqSpline.py:498-504 (no whitening Jacobian term)."""
    x = np.asarray(x, F32)
    xw = ((x - p.data_mean) / np.sqrt(np.diag(p.data_cov))).astype(F32)
    y, logdet = forward(p, xw)
    return (logdet + base_log_prob(p, y)).astype(F32)


def sample(p: FlowParams, key, n: int):
    """rqSpline.py:490-496: base.sample -> inverse -> un-whiten.  base cov must be I here."""
    z = rng.normal(key, (n, p.n_features))
    L = np.linalg.cholesky(p.base_cov.astype(np.float64)).astype(F32)
    z = (p.base_mean + z @ L.T).astype(F32)
    x, _ = inverse(p, z)
    return (x * np.sqrt(np.diag(p.data_cov)) + p.data_mean).astype(F32)
