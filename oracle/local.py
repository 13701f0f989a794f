"""Oracle: local kernels (MALA / HMC / Gaussian random walk) and TakeSerialSteps.  TEST ONLY.

Restates, vectorised over chains in numpy fp32:
  * MALA.kernel                  -- src/flowMC/resource/kernel/MALA.py:26-89
  * HMC.kernel / leapfrog_*      -- src/flowMC/resource/kernel/HMC.py:47-50,71-96,98-151
  * GaussianRandomWalk.kernel    -- src/flowMC/resource/kernel/Gaussian_random_walk.py:25-61
  * TakeSteps.__call__ / TakeSerialSteps.body,sample
                                 -- src/flowMC/strategy/take_steps.py:60-144,156-180
``jax.scipy.stats.multivariate_normal.logpdf`` with a scalar covariance (used by MALA,
MALA.py:76-81) is restated from jax 0.5.0:  -1/2 * (y.y)/cov - n/2 * (log(2 pi) + log(cov)).

Each kernel takes *per-chain* keys uint32[n,2] exactly like the vmapped reference does.
"""
from __future__ import annotations

import numpy as np

from . import rng
from .targets import TARGETS

F32 = np.float32
_LOG_2PI = F32(np.log(2 * np.pi))


def _mvn_logpdf_scalar_cov(x, mean, cov):
    y = (x - mean).astype(F32)
    n = x.shape[-1]
    yy = np.sum(y * y, axis=-1, dtype=F32)
    return (F32(-0.5) * yy / F32(cov) - F32(n / 2) * (_LOG_2PI + np.log(F32(cov)))).astype(F32)


def _split_pair(keys):
    """vmapped jax.random.split(key): keys [n,2] -> (k0 [n,2], k1 [n,2])."""
    s = rng.split(keys, 2)  # [n,2,2]
    return s[:, 0, :], s[:, 1, :]


def _log_uniform(keys):
    u = rng.uniform(keys, ())
    with np.errstate(divide="ignore"):
        return np.log(u).astype(F32)


def mala_kernel(keys, position, log_prob, target, data, step_size):
    """MALA.py:26-89.  The incoming log_prob is ignored by the reference (recomputed)."""
    tgt = TARGETS[target]
    x = position.astype(F32)
    n, d = x.shape
    key1, key2 = _split_pair(keys)
    dt = F32(step_size)
    dt2 = F32(dt * dt)
    lp0, g0 = tgt.logp_grad(x, data)
    z = rng.normal(key1, (d,))
    prop = (x + (dt2 * g0) / F32(2)).astype(F32)
    prop = (prop + dt * z).astype(F32)
    lp1, g1 = tgt.logp_grad(prop, data)
    ratio = (lp1 - lp0).astype(F32)
    ratio = ratio - _mvn_logpdf_scalar_cov(prop, (x + (dt2 * g0) / F32(2)).astype(F32), dt2)
    ratio = (ratio + _mvn_logpdf_scalar_cov(x, (prop + (dt2 * g1) / F32(2)).astype(F32), dt2)).astype(F32)
    log_u = _log_uniform(key2)
    acc = log_u < ratio
    new_x = np.where(acc[:, None], prop, x).astype(F32)
    new_lp = np.where(acc, lp1, lp0).astype(F32)
    return new_x, new_lp, acc, dict(ratio=ratio, log_u=log_u, z=z)


def grw_kernel(keys, position, log_prob, target, data, step_size):
    """Gaussian_random_walk.py:25-61."""
    tgt = TARGETS[target]
    x = position.astype(F32)
    n, d = x.shape
    key1, key2 = _split_pair(keys)
    z = rng.normal(key1, (d,))
    prop = (x + z * F32(step_size)).astype(F32)
    lp1, _ = tgt.logp_grad(prop, data)
    log_u = _log_uniform(key2)
    ratio = (lp1 - log_prob).astype(F32)
    acc = log_u < ratio
    new_x = np.where(acc[:, None], prop, x).astype(F32)
    new_lp = np.where(acc, lp1, log_prob).astype(F32)
    return new_x, new_lp, acc, dict(ratio=ratio, log_u=log_u, z=z)


def hmc_setup(condition_matrix, d):
    """Host-side constants of HMC.py:133-136,121: L = chol(inv(M)) and column sums of M."""
    M = np.asarray(condition_matrix, dtype=np.float64)
    if M.ndim == 0:
        raise ValueError("HMC condition_matrix must be a 2-D matrix (HMC.py:135 calls linalg.inv)")
    M32 = M.astype(F32)
    L = np.linalg.cholesky(np.linalg.inv(M32.astype(np.float64))).astype(F32)
    colsum = M32.sum(axis=0, dtype=F32).astype(F32)
    return L, colsum


def hmc_leapfrog(x, p, target, data, step_size, n_leapfrog, colsum):
    """HMC.py:71-96: n_leapfrog+2 iterations with coefficient rows [0,.5],[1,1]*n,[1,.5]."""
    tgt = TARGETS[target]
    eps = F32(step_size)
    coefs = np.ones((n_leapfrog + 2, 2), dtype=F32)
    coefs[0] = (0.0, 0.5)
    coefs[-1] = (1.0, 0.5)
    lp = None
    for i in range(n_leapfrog + 2):
        x = (x + eps * coefs[i, 0] * (p * colsum)).astype(F32)
        lp, g = tgt.logp_grad(x, data)
        p = (p - eps * coefs[i, 1] * (-g)).astype(F32)
    return x, p, lp


def hmc_kernel(keys, position, log_prob, target, data, step_size, n_leapfrog, L, colsum):
    """HMC.py:98-151."""
    x = position.astype(F32)
    n, d = x.shape
    key1, key2 = _split_pair(keys)
    z = rng.normal(key1, (d,))
    p = (z @ L.T).astype(F32)
    kin0 = (F32(0.5) * np.sum(p * p * colsum, axis=-1, dtype=F32)).astype(F32)
    H = (-log_prob + kin0).astype(F32)
    xp, pp, lp1 = hmc_leapfrog(x, p, target, data, step_size, n_leapfrog, colsum)
    pe = (-lp1).astype(F32)
    kin1 = (F32(0.5) * np.sum(pp * pp * colsum, axis=-1, dtype=F32)).astype(F32)
    ham = (pe + kin1).astype(F32)
    log_acc = (H - ham).astype(F32)
    log_u = _log_uniform(key2)
    acc = log_u < log_acc
    new_x = np.where(acc[:, None], xp, x).astype(F32)
    new_lp = np.where(acc, -pe, log_prob).astype(F32)
    return new_x, new_lp, acc, dict(ratio=log_acc, log_u=log_u, z=z)


def make_kernel(kind, **kw):
    if kind == "MALA":
        return lambda k, x, lp, t, dat: mala_kernel(k, x, lp, t, dat, kw["step_size"])
    if kind == "GRW":
        return lambda k, x, lp, t, dat: grw_kernel(k, x, lp, t, dat, kw["step_size"])
    if kind == "HMC":
        L, colsum = hmc_setup(kw["condition_matrix"], None)
        return lambda k, x, lp, t, dat: hmc_kernel(k, x, lp, t, dat, kw["step_size"], kw["n_leapfrog"], L, colsum)
    raise ValueError(kind)


def take_serial_steps(rng_key, initial_position, target, data, kernel, n_steps, thinning=1,
                      chain_offset=0, n_chains_total=None, return_debug=False):
    """take_steps.py:60-144 + 156-180 for one call.

    Returns (new_rng_key, positions[n, n_out, d], log_probs[n, n_out], accepts[n, n_out] (float32),
    last_position[n, d]) where n_out = len(range(0, n_steps, thinning)).
    ``chain_offset``/``n_chains_total`` select a shard of the global chain index space (the
    per-chain key is split(subkey, n_chains_total)[chain_offset + i], so any sharding gives
    identical chains).
    """
    x = np.asarray(initial_position, dtype=F32)
    n, d = x.shape
    n_tot = n if n_chains_total is None else n_chains_total
    ks = rng.split(rng_key, 2)
    new_key, subkey = ks[0], ks[1]
    chain_keys = rng.split(subkey, n_tot)[chain_offset:chain_offset + n]
    lp, _ = TARGETS[target].logp_grad(x, data)
    pos, lps, accs, dbg = [], [], [], []
    for t in range(n_steps):
        k0, k1 = _split_pair(chain_keys)
        chain_keys, sub = k0, k1
        x_in, lp_in = x, lp
        x, lp, acc, info = kernel(sub, x, lp, target, data)
        if return_debug:
            info.update(keys=sub.copy(), x_in=x_in.copy(), lp_in=lp_in.copy(), x_out=x.copy(), lp_out=lp.copy(),
                        acc=acc.copy())
        if t % thinning == 0:
            pos.append(x.copy())
            lps.append(lp.copy())
            accs.append(acc.astype(F32))
        if return_debug:
            dbg.append(info)
    positions = np.stack(pos, axis=1)
    out = (new_key, positions, np.stack(lps, axis=1), np.stack(accs, axis=1), positions[:, -1].copy())
    return out + (dbg,) if return_debug else out
