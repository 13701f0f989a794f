"""Oracle: target log-densities and analytic gradients, numpy fp32.  TEST ONLY.

The reference takes an arbitrary Python callable ``logpdf(x, data)`` and differentiates it
with ``jax.value_and_grad`` (src/flowMC/resource/logPDF.py:60-61, resource/kernel/MALA.py:59,
HMC.py:76-79).  The B200 path replaces that with registered device functions carrying an
analytic gradient, so the oracle restates each built-in target with the same closed forms.
Targets that come from the reference tree:
  * ``iso_gaussian``  -- test/unit/test_kernels.py:14-15 (c=0.5, mu=0),
                         test/unit/test_strategies.py:23-24 and test/integration/test_quickstart.py:7-8
                         (c=0.5, mu=data["data"]), test/unit/test_resources.py:14-15 (c=1).
  * ``dual_moon``     -- docs/tutorials/dualmoon.ipynb:77-84 (mu=0) and
                         test/integration/test_MALA.py:14-23 (mu=data["data"]).
Builder-defined targets for the BASELINE.json configs:
  ``ar1_gaussian`` (C2), ``dense_gaussian`` (C2 dense variant), ``rosenbrock`` (C3),
  ``gaussian_mixture`` (C4/C5).

Every function takes x float32[n, d] and the packed float32 ``data`` vector (same packing as
flowmc_b200/targets.py) and returns float32.  Gradients are checked against float64 central
differences in tests/test_oracle_targets.py.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def _f32(a):
    return np.asarray(a, dtype=F32)


def _lse2(a, b):
    m = np.maximum(a, b)
    return (m + np.log(np.exp(a - m) + np.exp(b - m))).astype(F32)


class IsoGaussian:
    name = "iso_gaussian"

    @staticmethod
    def pack(d, c=0.5, mu=None):
        mu = np.zeros(d, F32) if mu is None else _f32(mu).reshape(d)
        return np.concatenate([[F32(c)], mu]).astype(F32)

    @staticmethod
    def logp_grad(x, data):
        x = _f32(x)
        c, mu = data[0], data[1:]
        r = (x - mu).astype(F32)
        lp = (-c * np.sum(r * r, axis=-1, dtype=F32)).astype(F32)
        g = (F32(-2.0) * c * r).astype(F32)
        return lp, g


class DualMoon:
    name = "dual_moon"

    @staticmethod
    def pack(d, mu=None):
        return (np.zeros(d, F32) if mu is None else _f32(mu).reshape(d)).astype(F32)

    @staticmethod
    def logp_grad(x, data):
        x = _f32(x)
        mu = data
        r = (x - mu).astype(F32)
        nrm = np.sqrt(np.sum(r * r, axis=-1, dtype=F32)).astype(F32)
        t = ((nrm - F32(2.0)) / F32(0.1)).astype(F32)
        term1 = (F32(0.5) * t * t).astype(F32)
        a = ((x[:, 0] - F32(3.0)) / F32(0.8)).astype(F32)
        b = ((x[:, 0] + F32(3.0)) / F32(0.8)).astype(F32)
        t2a, t2b = F32(-0.5) * a * a, F32(-0.5) * b * b
        c = ((x[:, 1] - F32(3.0)) / F32(0.6)).astype(F32)
        e = ((x[:, 1] + F32(3.0)) / F32(0.6)).astype(F32)
        t3a, t3b = F32(-0.5) * c * c, F32(-0.5) * e * e
        l2, l3 = _lse2(t2a, t2b), _lse2(t3a, t3b)
        lp = (-(term1 - l2 - l3)).astype(F32)
        with np.errstate(divide="ignore", invalid="ignore"):
            g = (-(t / F32(0.1))[:, None] * (r / nrm[:, None])).astype(F32)
        wa, wb = np.exp(t2a - l2), np.exp(t2b - l2)
        g[:, 0] += (wa * (-a / F32(0.8)) + wb * (-b / F32(0.8))).astype(F32)
        wc, we = np.exp(t3a - l3), np.exp(t3b - l3)
        g[:, 1] += (wc * (-c / F32(0.6)) + we * (-e / F32(0.6))).astype(F32)
        return lp, g.astype(F32)


class AR1Gaussian:
    """Zero-mean Gaussian with covariance rho^|i-j| => tridiagonal precision."""
    name = "ar1_gaussian"

    @staticmethod
    def pack(d, rho=0.9):
        return np.array([rho], dtype=F32)

    @staticmethod
    def precision(d, rho):
        a = 1.0 / (1.0 - rho * rho)
        P = np.zeros((d, d))
        for j in range(d):
            P[j, j] = a * (1.0 + (rho * rho if 0 < j < d - 1 else 0.0))
            if j > 0:
                P[j, j - 1] = -a * rho
            if j < d - 1:
                P[j, j + 1] = -a * rho
        return P

    @staticmethod
    def logp_grad(x, data):
        x = _f32(x)
        n, d = x.shape
        rho = F32(data[0])
        a = F32(F32(1.0) / (F32(1.0) - rho * rho))
        diag = np.full(d, F32(1.0) + rho * rho, dtype=F32)
        diag[0] = F32(1.0)
        diag[-1] = F32(1.0)
        left = np.zeros_like(x)
        right = np.zeros_like(x)
        left[:, 1:] = x[:, :-1]
        right[:, :-1] = x[:, 1:]
        px = (a * (diag * x - rho * (left + right))).astype(F32)
        lp = (F32(-0.5) * np.sum(x * px, axis=-1, dtype=F32)).astype(F32)
        return lp, (-px).astype(F32)


class DenseGaussian:
    name = "dense_gaussian"

    @staticmethod
    def pack(d, precision=None):
        P = np.eye(d) if precision is None else np.asarray(precision)
        return _f32(P).reshape(d * d)

    @staticmethod
    def logp_grad(x, data):
        x = _f32(x)
        n, d = x.shape
        P = data.reshape(d, d)
        px = (x @ P.T).astype(F32)
        lp = (F32(-0.5) * np.sum(x * px, axis=-1, dtype=F32)).astype(F32)
        return lp, (-px).astype(F32)


class Rosenbrock:
    name = "rosenbrock"

    @staticmethod
    def pack(d):
        return np.zeros(1, F32)

    @staticmethod
    def logp_grad(x, data):
        x = _f32(x)
        xi, xn = x[:, :-1], x[:, 1:]
        u = (xn - xi * xi).astype(F32)
        v = (F32(1.0) - xi).astype(F32)
        lp = (-np.sum((F32(100.0) * u * u + v * v) / F32(20.0), axis=-1, dtype=F32)).astype(F32)
        g = np.zeros_like(x)
        g[:, :-1] += ((F32(400.0) * xi * u + F32(2.0) * v) / F32(20.0)).astype(F32)
        g[:, 1:] += ((F32(-200.0) * u) / F32(20.0)).astype(F32)
        return lp, g.astype(F32)


class GaussianMixture:
    """K isotropic components: logp = logsumexp_k(logw_k - 0.5*inv_var*|x-mu_k|^2)."""
    name = "gaussian_mixture"

    @staticmethod
    def pack(d, means, inv_var=1.0, logw=None):
        means = _f32(means)
        K = means.shape[0]
        assert means.shape == (K, d)
        logw = np.full(K, -np.log(K), F32) if logw is None else _f32(logw)
        return np.concatenate([[F32(K), F32(inv_var)], logw, means.reshape(-1)]).astype(F32)

    @staticmethod
    def logp_grad(x, data):
        x = _f32(x)
        n, d = x.shape
        K = int(data[0])
        iv = F32(data[1])
        logw = data[2:2 + K]
        mu = data[2 + K:2 + K + K * d].reshape(K, d)
        diff = (mu[None, :, :] - x[:, None, :]).astype(F32)  # n,K,d
        sq = np.sum(diff * diff, axis=-1, dtype=F32)
        e = (logw[None, :] - F32(0.5) * iv * sq).astype(F32)
        m = e.max(axis=-1, keepdims=True)
        p = np.exp(e - m).astype(F32)
        s = p.sum(axis=-1, keepdims=True, dtype=F32)
        lp = (m[:, 0] + np.log(s[:, 0])).astype(F32)
        w = (p / s).astype(F32)
        g = (iv * np.einsum("nk,nkd->nd", w, diff)).astype(F32)
        return lp, g


TARGETS = {t.name: t for t in (IsoGaussian, DualMoon, AR1Gaussian, DenseGaussian, Rosenbrock, GaussianMixture)}


def logp(name, x, data):
    return TARGETS[name].logp_grad(x, data)[0]


def logp_grad(name, x, data):
    return TARGETS[name].logp_grad(x, data)
