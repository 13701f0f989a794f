"""Test helpers: move flow parameters between the oracle's FlowParams and the device model."""
import numpy as np
import torch

from oracle import flow as oflow
from oracle import rng


def model_from_params(p):
    """Device MaskedCouplingRQSpline holding the oracle parameters ``p``."""
    from flowmc_b200.resource.model.nf_model.rqSpline import MaskedCouplingRQSpline
    m = MaskedCouplingRQSpline(p.n_features, p.n_layers, p.hidden, p.num_bins, None, (p.range_min, p.range_max))
    for l in range(p.n_layers):
        for i in range(len(p.W)):
            m.weight(l, i).copy_(torch.from_numpy(p.W[i][l]))
            m.bias(l, i).copy_(torch.from_numpy(p.b[i][l]))
        m.affine(l).copy_(torch.tensor([p.scale[l], p.shift[l]]))
    m.data_mean.copy_(torch.from_numpy(p.data_mean))
    m.data_cov.copy_(torch.from_numpy(p.data_cov))
    m.base_mean.copy_(torch.from_numpy(p.base_mean))
    m.base_cov.copy_(torch.from_numpy(p.base_cov))
    return m


def params_from_model(m):
    p = oflow.init_params(rng.PRNGKey(0), m.n_features, m.n_layers, m.hidden_size, m.num_bins, m.spline_range)
    for l in range(p.n_layers):
        for i in range(len(p.W)):
            p.W[i][l] = m.weight(l, i).cpu().numpy()
            p.b[i][l] = m.bias(l, i).cpu().numpy()
        a = m.affine(l).cpu().numpy()
        p.scale[l], p.shift[l] = a[0], a[1]
    p.data_mean = m.data_mean.cpu().numpy().copy()
    p.data_cov = m.data_cov.cpu().numpy().copy()
    p.base_mean = m.base_mean.cpu().numpy().copy()
    p.base_cov = m.base_cov.cpu().numpy().copy()
    return p


def random_params(seed, d, n_layers, hidden, num_bins, gain=3.0, affine=0.2, whiten=True):
    """Oracle parameters away from the initialisation: larger conditioner weights (so the spline
    knots are far from uniform), non-zero ScalarAffine, non-trivial whitening statistics."""
    r = np.random.default_rng(seed)
    p = oflow.init_params(rng.PRNGKey(seed), d, n_layers, hidden, num_bins)
    for i in range(len(p.W)):
        p.W[i] = (p.W[i] * np.float32(gain)).astype(np.float32)
        p.b[i] = (p.b[i] * np.float32(gain)).astype(np.float32)
    p.scale = (affine * r.standard_normal(n_layers)).astype(np.float32)
    p.shift = (affine * r.standard_normal(n_layers)).astype(np.float32)
    if whiten:
        p.data_mean = r.standard_normal(d).astype(np.float32)
        a = r.standard_normal((d, d)) * 0.3
        p.data_cov = (a @ a.T + np.diag(0.5 + r.random(d))).astype(np.float32)
    return p


def nvp_model_from_params(p):
    """Device RealNVP holding the oracle parameters ``p`` (oracle.realnvp.NVPParams)."""
    from flowmc_b200.resource.model.nf_model.realNVP import RealNVP
    m = RealNVP(p.n_features, p.n_layers, p.n_hidden, None)
    for l in range(p.n_layers):
        for w in range(2):
            m.W1(l, w).copy_(torch.from_numpy(p.W1[w, l]))
            m.b1(l, w).copy_(torch.from_numpy(p.b1[w, l]))
            m.W2(l, w).copy_(torch.from_numpy(p.W2[w, l]))
            m.b2(l, w).copy_(torch.from_numpy(p.b2[w, l]))
        m.layer_mask(l).copy_(torch.from_numpy(p.mask[l]))
    m.data_mean.copy_(torch.from_numpy(p.data_mean))
    m.data_cov.copy_(torch.from_numpy(p.data_cov))
    m.base_mean.copy_(torch.from_numpy(p.base_mean))
    m.base_cov.copy_(torch.from_numpy(p.base_cov))
    return m


def nvp_params_from_model(m):
    from oracle import realnvp as onvp
    p = onvp.init_params(rng.PRNGKey(0), m.n_features, m.n_layers, m.n_hidden)
    for l in range(p.n_layers):
        for w in range(2):
            p.W1[w, l] = m.W1(l, w).cpu().numpy()
            p.b1[w, l] = m.b1(l, w).cpu().numpy()
            p.W2[w, l] = m.W2(l, w).cpu().numpy()
            p.b2[w, l] = m.b2(l, w).cpu().numpy()
        p.mask[l] = m.layer_mask(l).cpu().numpy()
    p.data_mean = m.data_mean.cpu().numpy().copy()
    p.data_cov = m.data_cov.cpu().numpy().copy()
    p.base_mean = m.base_mean.cpu().numpy().copy()
    p.base_cov = m.base_cov.cpu().numpy().copy()
    return p
