"""Oracle: NFProposal global steps, flow training (loss/grad, clip+AdamW, epochs), TrainModel data
selection.  TEST ONLY.

Restates:
  * NFProposal.kernel / sample_flow            src/flowMC/resource/kernel/NF_proposal.py:27-172
  * TakeGroupSteps.sample + TakeSteps.__call__ src/flowMC/strategy/take_steps.py:60-144,191-206
  * NFModel.loss_fn / train_step / train_epoch / train   src/flowMC/resource/model/nf_model/base.py:98-210
  * Optimizer (clip_by_global_norm(1.0) -> adamw(lr, b1=momentum, wd=1e-4))   src/flowMC/resource/optimizer.py:19-23
    (optax 0.2.4 semantics, restated from its published source)
  * TrainModel.__call__ data selection          src/flowMC/strategy/train_model.py:47-112
Gradients come from torch autograd over a torch restatement of oracle/flow.py (float64 by default,
so they double as a check of the hand-written CUDA backward).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import flow as oflow
from . import realnvp as orealnvp
from . import rng
from .targets import TARGETS

F32 = np.float32


# ------------------------------------------------------------------------------ NFProposal
def _model_log_prob(p, x):
    """model.log_prob for either flow family (FlowParams: spline flow, NVPParams: RealNVP)."""
    if isinstance(p, orealnvp.NVPParams):
        return orealnvp.log_prob(p, x)
    return oflow.log_prob(p, x)


def sample_flow(p, keys, n_steps, n_batch_size):
    """NF_proposal.py:130-172 for a batch of per-chain keys [n,2] -> (positions [n,S,d], flow log-probs [n,S])."""
    n = keys.shape[0]
    d = p.n_features
    if n_steps > n_batch_size:
        n_batch = math.ceil(n_steps / n_batch_size)
        n_sample = math.ceil(n_steps / n_batch)
        ks = keys
        pos, lps = [], []
        for _ in range(n_batch):
            s = rng.split(ks, 2)
            ks, sub = s[:, 0], s[:, 1]
            z = rng.normal(sub, (n_sample, d))                       # [n, n_sample, d]
            x = _sample_from_z(p, z.reshape(-1, d))
            pos.append(x.reshape(n, n_sample, d))
            lps.append(_model_log_prob(p, x).reshape(n, n_sample))
        pos = np.concatenate(pos, axis=1)[:, :n_steps]
        lps = np.concatenate(lps, axis=1)[:, :n_steps]
    else:
        z = rng.normal(keys, (n_steps, d))
        x = _sample_from_z(p, z.reshape(-1, d))
        pos = x.reshape(n, n_steps, d)
        lps = _model_log_prob(p, x).reshape(n, n_steps)
    return pos.astype(F32), lps.astype(F32)


def _sample_from_z(p, z):
    if isinstance(p, orealnvp.NVPParams):
        return orealnvp.sample_from_z(p, z)
    L = np.linalg.cholesky(p.base_cov.astype(np.float64)).astype(F32)
    z = (p.base_mean + z @ L.T).astype(F32)
    x, _ = oflow.inverse(p, z)
    return (x * np.sqrt(np.diag(p.data_cov)) + p.data_mean).astype(F32)


def nf_proposal_kernel(p, keys, position, log_prob, target, data, n_steps, n_batch_size):
    """NF_proposal.py:27-128, vectorised over chains (each chain has its own key)."""
    tgt = TARGETS[target]
    n, d = position.shape
    s = rng.split(keys, 2)
    rk, sub = s[:, 0], s[:, 1]
    lp_nf_cur = _model_log_prob(p, position)
    prop, lp_nf_prop = sample_flow(p, sub, n_steps, n_batch_size)
    lp_prop = tgt.logp_grad(prop.reshape(-1, d), data)[0].reshape(n, n_steps)
    x = position.astype(F32).copy()
    lp = log_prob.astype(F32).copy()
    out_x, out_lp, out_acc, dbg = [], [], [], []
    for t in range(n_steps):
        s = rng.split(rk, 2)
        rk, sk = s[:, 0], s[:, 1]
        ratio = ((lp_prop[:, t] - lp) - (lp_nf_prop[:, t] - lp_nf_cur)).astype(F32)
        with np.errstate(divide="ignore"):
            log_u = np.log(rng.uniform(sk, ())).astype(F32)
        acc = log_u < ratio
        x = np.where(acc[:, None], prop[:, t], x).astype(F32)
        lp = np.where(acc, lp_prop[:, t], lp).astype(F32)
        lp_nf_cur = np.where(acc, lp_nf_prop[:, t], lp_nf_cur).astype(F32)
        out_x.append(x.copy()); out_lp.append(lp.copy()); out_acc.append(acc.copy())
        dbg.append(dict(ratio=ratio, log_u=log_u))
    return np.stack(out_x, 1), np.stack(out_lp, 1), np.stack(out_acc, 1), dict(steps=dbg, proposals=prop,
                                                                                   lp_nf_prop=lp_nf_prop, lp_prop=lp_prop)


def take_group_steps(rng_key, initial_position, p, target, data, n_steps, n_batch_size, thinning=1,
                     chain_offset=0, n_chains_total=None):
    """TakeSteps.__call__ with TakeGroupSteps.sample (take_steps.py:60-144,191-206)."""
    x0 = np.asarray(initial_position, F32)
    n, d = x0.shape
    n_tot = n if n_chains_total is None else n_chains_total
    ks = rng.split(rng_key, 2)
    new_key, subkey = ks[0], ks[1]
    chain_keys = rng.split(subkey, n_tot)[chain_offset:chain_offset + n]
    lp0 = TARGETS[target].logp_grad(x0, data)[0]
    pos, lp, acc, dbg = nf_proposal_kernel(p, chain_keys, x0, lp0, target, data, n_steps, n_batch_size)
    pos, lp, acc = pos[:, ::thinning], lp[:, ::thinning], acc[:, ::thinning].astype(F32)
    return new_key, pos, lp, acc, pos[:, -1].copy(), dbg


# ------------------------------------------------------------------------------ torch flow (autograd)
PARAM_ORDER_NOTE = "flat order: per layer [W_0,b_0,...,W_last,b_last,scale,shift], then data_mean, data_cov, base_mean, base_cov"


def to_torch(p, dtype=torch.float64, requires_grad=True):
    tp = dict(W=[torch.tensor(w, dtype=dtype, requires_grad=requires_grad) for w in p.W],
              b=[torch.tensor(b, dtype=dtype, requires_grad=requires_grad) for b in p.b],
              scale=torch.tensor(p.scale, dtype=dtype, requires_grad=requires_grad),
              shift=torch.tensor(p.shift, dtype=dtype, requires_grad=requires_grad))
    return tp


def _t_get_params(p, raw, dtype):
    K = p.num_bins
    n = raw.shape[0]
    r = raw.reshape(n, p.n_features, 3 * K + 1)
    uw, uh, us = r[..., :K], r[..., K:2 * K], r[..., 2 * K:]
    rmin, rmax = p.range_min, p.range_max
    sc = float(F32(F32(rmax - rmin) - F32(K) * oflow.MIN_BIN))
    mb = float(oflow.MIN_BIN)
    bw = torch.softmax(uw, -1) * sc + mb
    bh = torch.softmax(uh, -1) * sc + mb
    lo = torch.full(r.shape[:-1] + (1,), rmin, dtype=dtype)
    hi = torch.full(r.shape[:-1] + (1,), rmax, dtype=dtype)
    x_pos = torch.cat([lo, rmin + torch.cumsum(bw[..., :-1], -1), hi], -1)
    y_pos = torch.cat([lo, rmin + torch.cumsum(bh[..., :-1], -1), hi], -1)
    offset = float(F32(np.log(np.exp(F32(1.0) - oflow.MIN_SLOPE) - F32(1.0))))
    slopes = torch.nn.functional.softplus(us + offset) + float(oflow.MIN_SLOPE)
    return x_pos, y_pos, slopes


def _t_spline_fwd(x, x_pos, y_pos, slopes):
    below = x <= x_pos[..., 0]
    above = x >= x_pos[..., -1]
    inb = (x[..., None] >= x_pos[..., :-1]) & (x[..., None] < x_pos[..., 1:])
    none = ~inb.any(-1)
    first = torch.zeros_like(inb)
    first[..., 0] = True
    inb = torch.where(none[..., None], first, inb).to(x.dtype)
    pick = lambda a: ((inb * a[..., :-1]).sum(-1), (inb * a[..., 1:]).sum(-1))
    xl, xr = pick(x_pos); yl, yr = pick(y_pos); dl, dr = pick(slopes)
    bw, bh = xr - xl, yr - yl
    s = bh / bw
    z = torch.clamp((x - xl) / bw, 0.0, 1.0)
    sq_z = z * z
    z1mz = z - sq_z
    sq_1mz = (1.0 - z) ** 2
    st = dr + dl - 2.0 * s
    num = bh * (s * sq_z + dl * z1mz)
    den = s + st * z1mz
    y = yl + num / den
    logdet = 2.0 * torch.log(s) + torch.log(dr * sq_z + 2.0 * s * z1mz + dl * sq_1mz) - 2.0 * torch.log(den)
    y = torch.where(below, (x - x_pos[..., 0]) * slopes[..., 0] + y_pos[..., 0], y)
    y = torch.where(above, (x - x_pos[..., -1]) * slopes[..., -1] + y_pos[..., -1], y)
    logdet = torch.where(below, torch.log(slopes[..., 0]), logdet)
    logdet = torch.where(above, torch.log(slopes[..., -1]), logdet)
    return y, logdet


def torch_log_prob(p, tp, x, dtype=torch.float64):
    """log_prob of oracle/flow.py in torch (differentiable in tp)."""
    d = p.n_features
    x = torch.as_tensor(np.asarray(x), dtype=dtype)
    mean = torch.tensor(p.data_mean, dtype=dtype)
    std = torch.sqrt(torch.tensor(np.diag(p.data_cov).copy(), dtype=dtype))
    x = (x - mean) / std
    logdet = torch.zeros(x.shape[0], dtype=dtype)
    n_lin = len(tp["W"])
    for l in range(p.n_layers):
        x = (x + tp["shift"][l]) * torch.exp(tp["scale"][l])
        logdet = logdet + d * tp["scale"][l]
        m = torch.tensor(p.mask(l), dtype=dtype)
        h = x * m
        for i in range(n_lin):
            h = h @ tp["W"][i][l].T + tp["b"][i][l]
            if i < n_lin - 1:
                h = torch.tanh(h)
        xp, yp, sl = _t_get_params(p, h, dtype)
        t, ld = _t_spline_fwd(x, xp, yp, sl)
        x = (1 - m) * t + m * x
        logdet = logdet + ((1 - m) * ld).sum(-1)
    bc = torch.tensor(np.diag(p.base_cov).copy(), dtype=dtype)
    bm = torch.tensor(p.base_mean, dtype=dtype)
    base = -0.5 * (((x - bm) ** 2) / bc).sum(-1) - d / 2 * math.log(2 * math.pi) - 0.5 * torch.log(bc).sum()
    return logdet + base


def loss_and_grads(p, x, dtype=torch.float64):
    """NFModel.loss_fn (base.py:98-100): -mean(log_prob); grads as a FlowParams-shaped dict of numpy arrays."""
    tp = to_torch(p, dtype)
    loss = -torch_log_prob(p, tp, x, dtype).mean()
    leaves = tp["W"] + tp["b"] + [tp["scale"], tp["shift"]]
    gs = torch.autograd.grad(loss, leaves)
    nW = len(tp["W"])
    g = dict(W=[t.numpy().astype(F32) for t in gs[:nW]], b=[t.numpy().astype(F32) for t in gs[nW:2 * nW]],
             scale=gs[2 * nW].numpy().astype(F32), shift=gs[2 * nW + 1].numpy().astype(F32))
    return float(loss.detach()), g


# ------------------------------------------------------------------------------ flat layout + optimizer
def flatten(p_or_g, p=None):
    """Flat float32 vector in the device blob order (include/flowmc_b200.h FlowmcFlowDesc, without its alignment padding).  For a grads dict,
    the non-trainable tail (data_mean, data_cov, base_mean, base_cov) is zero."""
    if isinstance(p_or_g, dict):
        g, ref = p_or_g, p
        tail = [np.zeros_like(ref.data_mean), np.zeros_like(ref.data_cov), np.zeros_like(ref.base_mean),
                np.zeros_like(ref.base_cov)]
        W, b, scale, shift, L = g["W"], g["b"], g["scale"], g["shift"], ref.n_layers
    else:
        ref = p_or_g
        tail = [ref.data_mean, ref.data_cov, ref.base_mean, ref.base_cov]
        W, b, scale, shift, L = ref.W, ref.b, ref.scale, ref.shift, ref.n_layers
    parts = []
    for l in range(L):
        for i in range(len(W)):
            parts += [W[i][l].reshape(-1), b[i][l].reshape(-1)]
        parts += [scale[l:l + 1], shift[l:l + 1]]
    parts += [t.reshape(-1) for t in tail]
    return np.concatenate(parts).astype(F32)


def unflatten(p, flat):
    """Inverse of flatten for a FlowParams (returns a new FlowParams)."""
    q = p.copy()
    o = 0
    for l in range(p.n_layers):
        for i in range(len(p.W)):
            n = p.W[i][l].size
            q.W[i][l] = flat[o:o + n].reshape(p.W[i][l].shape); o += n
            n = p.b[i][l].size
            q.b[i][l] = flat[o:o + n]; o += n
        q.scale[l] = flat[o]; o += 1
        q.shift[l] = flat[o]; o += 1
    d = p.n_features
    q.data_mean = flat[o:o + d].copy(); o += d
    q.data_cov = flat[o:o + d * d].reshape(d, d).copy(); o += d * d
    q.base_mean = flat[o:o + d].copy(); o += d
    q.base_cov = flat[o:o + d * d].reshape(d, d).copy(); o += d * d
    assert o == flat.size
    return q


class AdamWState:
    def __init__(self, n):
        self.mu = np.zeros(n, F32)
        self.nu = np.zeros(n, F32)
        self.count = 0

    def copy(self):
        s = AdamWState(self.mu.size)
        s.mu, s.nu, s.count = self.mu.copy(), self.nu.copy(), self.count
        return s


def clip_adamw(params, grads, st: AdamWState, lr, b1=0.9, b2=0.999, eps=1e-8, wd=1e-4, max_norm=1.0):
    """optax.chain(clip_by_global_norm(1.0), adamw(lr, b1=momentum)) on flat fp32 vectors; returns new params."""
    g = grads.astype(F32)
    gnorm = F32(np.sqrt(np.sum(g.astype(F32) * g, dtype=F32)))
    if not (gnorm < F32(max_norm)):
        g = ((g / gnorm) * F32(max_norm)).astype(F32)
    st.mu = (F32(1 - b1) * g + F32(b1) * st.mu).astype(F32)
    st.nu = (F32(1 - b2) * (g * g) + F32(b2) * st.nu).astype(F32)
    st.count += 1
    bc1 = F32(1) - F32(b1) ** F32(st.count)
    bc2 = F32(1) - F32(b2) ** F32(st.count)
    mu_hat = (st.mu / bc1).astype(F32)
    nu_hat = (st.nu / bc2).astype(F32)
    u = (mu_hat / (np.sqrt(nu_hat) + F32(eps))).astype(F32)
    u = (u + F32(wd) * params).astype(F32)
    u = (F32(-lr) * u).astype(F32)
    return (params + u).astype(F32), float(gnorm)


def train(p, rng_key, data, st: AdamWState, lr, num_epochs, batch_size, momentum=0.9, grad_dtype=torch.float64):
    """NFModel.train (base.py:153-210) with the Optimizer's chain.  Returns (rng, best params, best state,
    losses); ``st`` is left untouched (the reference's optimiser state is functional)."""
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
            value, g = loss_and_grads(q, xb, grad_dtype)
            flat, _ = clip_adamw(flatten(q), flatten(g, q), st, lr, b1=momentum)
            q = unflatten(q, flat)
        losses[e] = value
        if losses[e] < best_loss:
            best, best_st, best_loss = q.copy(), st.copy(), losses[e]
    return key, best, best_st, losses


def select_training_data(rng_key, buffer, n_max_examples, history_window):
    """TrainModel.__call__ (train_model.py:66-81): finite rows, last `history_window` steps, choice with
    replacement.  Returns (rng_key after both splits' first halves, train subkey, data)."""
    n_chains, _, d = buffer.shape
    finite = np.isfinite(buffer).all(axis=-1)
    rows = buffer[finite].reshape(n_chains, -1, d)
    rows = rows[:, -history_window:].reshape(-1, d)
    ks = rng.split(rng_key, 2)
    key, sub = ks[0], ks[1]
    idx = rng.choice_with_replacement(sub, rows.shape[0], n_max_examples)
    out = rows[idx]
    ks = rng.split(key, 2)
    key, train_key = ks[0], ks[1]
    return key, train_key, out.astype(F32), idx
