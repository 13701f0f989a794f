"""MaskedCouplingRQSpline (reference: src/flowMC/resource/model/nf_model/rqSpline.py:360-504).

Same constructor and methods as the reference.  The model is ONE flat float32 device vector
(``self.params``, layout described by ``FlowmcFlowDesc`` in include/flowmc_b200.h) so that the
fused AdamW and the data-parallel gradient all-reduce each see a single buffer; ``forward`` /
``inverse`` / ``log_prob`` / ``sample`` are one C-ABI call each into the fused sm_100a kernels
(csrc/flow.cu).  Methods take batches ``[n, d]`` (the reference's per-sample methods are always
used under ``vmap``); a single sample ``[d]`` is accepted and returns un-batched results.

Initialisation reproduces the reference's key schedule and distributions bit for bit
(rqSpline.py:427-443, common.py:83-107, equinox.nn.Linear's uniform(+-1/sqrt(in)) init) using
the library's jax.random-compatible device generators.
"""
from __future__ import annotations

import ctypes as C
import json
import math
import os

import numpy as np
import torch

from .... import random as frandom
from ...._lib import FlowDesc, check, lib
from .base import NFModel

_u32p = C.POINTER(C.c_uint32)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def make_desc(n_features: int, n_layers: int, hidden_size, num_bins: int, spline_range) -> FlowDesc:
    desc = FlowDesc()
    hid = (C.c_int * len(hidden_size))(*[int(h) for h in hidden_size])
    check(lib.flowmc_flow_desc_init(C.byref(desc), int(n_features), int(n_layers), len(hidden_size), hid,
                                    int(num_bins), float(spline_range[0]), float(spline_range[1])))
    return desc


class MaskedCouplingRQSpline(NFModel):
    """Rational quadratic spline normalizing flow (masked coupling, MLP conditioner)."""

    def __repr__(self):
        return "MaskedCouplingRQSpline with n_features=" + str(self._n_features) + ", n_layers=" + str(self.n_layers)

    def __init__(self, n_features: int, n_layers: int, hidden_size: list, num_bins: int, key,
                 spline_range: tuple = (-10.0, 10.0), device=None, **kwargs):
        if kwargs.get("base_dist") is not None:
            raise NotImplementedError("flowmc_b200 supports the default Gaussian base distribution")
        if not torch.cuda.is_available():
            raise RuntimeError("flowmc_b200 needs a CUDA device (there is no CPU fallback)")
        self._n_features = int(n_features)
        self.n_layers = int(n_layers)
        self.hidden_size = [int(h) for h in hidden_size]
        self.num_bins = int(num_bins)
        self.spline_range = (float(spline_range[0]), float(spline_range[1]))
        self.desc = make_desc(n_features, n_layers, self.hidden_size, num_bins, self.spline_range)
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.params = torch.zeros(int(self.desc.n_params), dtype=torch.float32, device=dev)
        d = self._n_features
        self._view(self.desc.off_data_cov, (d, d)).copy_(torch.eye(d, device=dev))
        self._view(self.desc.off_base_cov, (d, d)).copy_(torch.eye(d, device=dev))
        if kwargs.get("data_mean") is not None:
            self._view(self.desc.off_data_mean, (d,)).copy_(torch.as_tensor(kwargs["data_mean"], dtype=torch.float32))
        if kwargs.get("data_cov") is not None:
            self._view(self.desc.off_data_cov, (d, d)).copy_(
                torch.atleast_2d(torch.as_tensor(kwargs["data_cov"], dtype=torch.float32)))
        if key is not None:
            self._init_weights(np.asarray(key, dtype=np.uint32))
        # execution path of the conditioner GEMMs: 3 = tcgen05 3xTF32 (default where the shape allows; fp32-grade),
        # 1 = tcgen05 plain TF32 (fast, ~1e-3 relative), 0 = fp32 CUDA cores.  FLOWMC_FLOW_TC overrides the default.
        self.tc_terms = int(os.environ.get("FLOWMC_FLOW_TC", "3"))
        self._tc_image = None

    # ---- parameter blob views ---------------------------------------------------------------
    def _view(self, off: int, shape, layer: int | None = None) -> torch.Tensor:
        o = int(off) + (0 if layer is None else layer * int(self.desc.layer_stride))
        return self.params[o:o + math.prod(shape)].view(*shape)

    @property
    def dims(self) -> list:
        return [int(self.desc.dims[i]) for i in range(self.desc.n_linear + 1)]

    def weight(self, layer: int, i: int) -> torch.Tensor:
        """W_i of coupling layer ``layer``: [out_i, in_i] (equinox Linear layout)."""
        dm = self.dims
        return self._view(self.desc.off_W[i], (dm[i + 1], dm[i]), layer)

    def bias(self, layer: int, i: int) -> torch.Tensor:
        return self._view(self.desc.off_b[i], (self.dims[i + 1],), layer)

    def affine(self, layer: int) -> torch.Tensor:
        """[scale, shift] of the layer's ScalarAffine (trainable, common.py:211-217)."""
        return self._view(self.desc.off_scale, (2,), layer)

    @property
    def data_mean(self) -> torch.Tensor:
        return self._view(self.desc.off_data_mean, (self._n_features,))

    @property
    def data_cov(self) -> torch.Tensor:
        return self._view(self.desc.off_data_cov, (self._n_features, self._n_features))

    @property
    def base_mean(self) -> torch.Tensor:
        return self._view(self.desc.off_base_mean, (self._n_features,))

    @property
    def base_cov(self) -> torch.Tensor:
        return self._view(self.desc.off_base_cov, (self._n_features, self._n_features))

    def _init_weights(self, key: np.ndarray):
        dm = self.dims
        n_lin = len(dm) - 1
        dev = self.params.device
        keys = frandom.split(key, self.n_layers)                       # rqSpline.py:442
        for l in range(self.n_layers):
            k = keys[l]
            for i in range(n_lin - 1):                                 # common.py:93-103
                k, sub1, sub2 = frandom.split(k, 3)
                _, bkey = frandom.split(sub1, 2)                       # eqx.nn.Linear: wkey, bkey = split(key, 2)
                lim = 1.0 / math.sqrt(dm[i])
                self.bias(l, i).copy_(frandom.uniform(bkey, (dm[i + 1],), -lim, lim, device=dev))
                w = frandom.normal(sub2, (dm[i + 1], dm[i]), device=dev)
                std = float(np.sqrt(np.float32(1e-2 / dm[i])))         # jnp.sqrt(scale / shape[i]) in float32
                self.weight(l, i).copy_(w * std)
            k, sub = frandom.split(k, 2)                               # common.py:104-107
            wkey, bkey = frandom.split(sub, 2)
            lim = 1.0 / math.sqrt(dm[-2])
            self.weight(l, n_lin - 1).copy_(frandom.uniform(wkey, (dm[-1], dm[-2]), -lim, lim, device=dev))
            self.bias(l, n_lin - 1).copy_(frandom.uniform(bkey, (dm[-1],), -lim, lim, device=dev))

    # ---- tensor-core weight image -------------------------------------------------------------
    def tc_supported(self) -> bool:
        return int(lib.flowmc_flow_tc_image_bytes(C.byref(self.desc))) > 0

    def prepare(self):
        """Refresh the packed tf32 hi/lo weight image the tcgen05 kernels stream (call after params change;
        every public method does).  Falls back to the fp32 CUDA-core kernels -- same results within the
        parity tolerance -- when the shape is outside the tensor-core path's limits or tc_terms == 0."""
        nbytes = int(lib.flowmc_flow_tc_image_bytes(C.byref(self.desc))) if self.tc_terms else 0
        if nbytes == 0:
            self.desc.tc_image, self.desc.tc_terms = None, 0
            return
        if self._tc_image is None or self._tc_image.numel() != nbytes:
            self._tc_image = torch.empty(nbytes, dtype=torch.uint8, device=self.params.device)
        with torch.cuda.device(self.params.device):
            check(lib.flowmc_flow_tc_pack(C.byref(self.desc), self.params.data_ptr(), self._tc_image.data_ptr(),
                                          _stream()))
        self.desc.tc_image, self.desc.tc_terms = self._tc_image.data_ptr(), int(self.tc_terms)

    def _prepare_bound(self, stream):
        """``prepare`` for a loop of training steps: the image exists and the shape cannot change, only the weights
        do -- one pack call per step, everything else resolved here."""
        self.prepare()
        if not self.desc.tc_terms:
            return lambda: None
        desc, params_ptr, image_ptr = C.byref(self.desc), self.params.data_ptr(), self._tc_image.data_ptr()
        pack = lib.flowmc_flow_tc_pack
        return lambda: check(pack(desc, params_ptr, image_ptr, stream))

    # ---- bijection ---------------------------------------------------------------------------
    def _prep(self, x):
        x = torch.as_tensor(x, dtype=torch.float32)
        if not x.is_cuda:
            x = x.to(self.params.device)
        single = x.dim() == 1
        x2 = (x.reshape(1, -1) if single else x.reshape(-1, x.shape[-1])).contiguous()
        if x2.shape[1] != self._n_features:
            raise ValueError(f"expected {self._n_features} features, got {x2.shape[1]}")
        return x2, single, x.shape[:-1]

    def _transform(self, fn, x):
        x2, single, lead = self._prep(x)
        self.prepare()
        n = x2.shape[0]
        y = torch.empty_like(x2)
        ld = torch.empty(n, dtype=torch.float32, device=x2.device)
        with torch.cuda.device(x2.device):
            check(fn(C.byref(self.desc), self.params.data_ptr(), x2.data_ptr(), n, y.data_ptr(), ld.data_ptr(),
                     _stream()))
        if single:
            return y[0], ld[0]
        return y.reshape(*lead, x2.shape[-1]), ld.reshape(*lead)   # (an empty batch cannot infer -1)

    def forward(self, x, key=None, condition=None):
        """Data -> latent (no whitening): returns (y, log_det) (rqSpline.py:450-468)."""
        return self._transform(lib.flowmc_flow_forward, x)

    def inverse(self, x, condition=None):
        """Latent -> data (rqSpline.py:470-488)."""
        return self._transform(lib.flowmc_flow_inverse, x)

    def __call__(self, x):
        return self.forward(x)

    def log_prob(self, x):
        """rqSpline.py:498-504."""
        x2, single, lead = self._prep(x)
        self.prepare()
        n = x2.shape[0]
        lp = torch.empty(n, dtype=torch.float32, device=x2.device)
        with torch.cuda.device(x2.device):
            check(lib.flowmc_flow_log_prob(C.byref(self.desc), self.params.data_ptr(), x2.data_ptr(), n,
                                           lp.data_ptr(), None, _stream()))
        return lp[0] if single else lp.reshape(*lead)

    def sample(self, rng_key, n_samples: int):
        """rqSpline.py:490-496: base.sample(key, n) -> inverse -> un-whiten."""
        key = np.ascontiguousarray(rng_key, dtype=np.uint32)
        self.prepare()
        out = torch.empty((int(n_samples), self._n_features), dtype=torch.float32, device=self.params.device)
        if int(n_samples) == 0:
            return out
        with torch.cuda.device(out.device):
            check(lib.flowmc_flow_sample(C.byref(self.desc), self.params.data_ptr(), None,
                                         key.ctypes.data_as(_u32p), int(n_samples), int(n_samples),
                                         out.data_ptr(), _stream()))
        return out

    # ---- resource ----------------------------------------------------------------------------
    def print_parameters(self):
        print("MaskedCouplingRQSpline parameters:")
        print(f"n_features: {self._n_features}, n_layers: {self.n_layers}, hidden_size: {self.hidden_size}, "
              f"num_bins: {self.num_bins}, spline_range: {self.spline_range}")

    def _header(self) -> dict:
        return dict(n_features=self._n_features, n_layers=self.n_layers, hidden_size=self.hidden_size,
                    num_bins=self.num_bins, spline_range=list(self.spline_range), n_params=int(self.desc.n_params))

    def _eqx_leaves(self) -> dict:
        from .... import eqx_io
        n_lin = len(self.dims) - 1
        W = [np.stack([self.weight(l, i).detach().cpu().numpy() for l in range(self.n_layers)]) for i in range(n_lin)]
        b = [np.stack([self.bias(l, i).detach().cpu().numpy() for l in range(self.n_layers)]) for i in range(n_lin)]
        aff = np.stack([self.affine(l).detach().cpu().numpy() for l in range(self.n_layers)])
        return eqx_io.leaves_from_arrays(self._n_features, self.n_layers, self.hidden_size, self.num_bins,
                                         self.spline_range, self.data_mean.cpu().numpy(), self.data_cov.cpu().numpy(),
                                         self.base_mean.cpu().numpy(), self.base_cov.cpu().numpy(), aff[:, 0], aff[:, 1],
                                         W, b)

    def save_model(self, path: str):
        """``path + ".eqx"`` in the reference's on-disk format (``eqx.tree_serialise_leaves``,
        nf_model/base.py:92-93; restated in flowmc_b200/eqx_io.py), so that the file loads into a real flowMC
        ``MaskedCouplingRQSpline`` of the same architecture and vice versa."""
        from .... import eqx_io
        with open(path + ".eqx", "wb") as f:
            eqx_io.write_eqx(f, self._n_features, self.n_layers, self.hidden_size, self.num_bins, self._eqx_leaves())

    def load_model(self, path: str) -> "MaskedCouplingRQSpline":
        """Returns a NEW model with this model's architecture and the weights of ``path + ".eqx"``
        (nf_model/base.py:95-96: ``eqx.tree_deserialise_leaves(path + ".eqx", self)``)."""
        from .... import eqx_io
        with open(path + ".eqx", "rb") as f:
            lv = eqx_io.read_eqx(f, self._n_features, self.n_layers, self.hidden_size, self.num_bins)
        m = MaskedCouplingRQSpline(self._n_features, self.n_layers, self.hidden_size, self.num_bins, None,
                                   self.spline_range, device=self.params.device)
        m.tc_terms = self.tc_terms
        dev = self.params.device
        m.data_mean.copy_(torch.from_numpy(lv["_data_mean"]).to(dev))
        m.data_cov.copy_(torch.from_numpy(lv["_data_cov"]).to(dev))
        m.base_mean.copy_(torch.from_numpy(lv["base_dist._mean"]).to(dev))
        m.base_cov.copy_(torch.from_numpy(lv["base_dist._cov"]).to(dev))
        for l in range(self.n_layers):
            m.affine(l).copy_(torch.tensor([lv["layers[0].bijector.scale"][l], lv["layers[0].bijector.shift"][l]]))
            for i in range(len(self.dims) - 1):
                m.weight(l, i).copy_(torch.from_numpy(lv[f"layers[1].bijector.conditioner.layers[{2 * i}].weight"][l]).to(dev))
                m.bias(l, i).copy_(torch.from_numpy(lv[f"layers[1].bijector.conditioner.layers[{2 * i}].bias"][l]).to(dev))
        return m

    save_resource = save_model
    load_resource = load_model

    def clone(self) -> "MaskedCouplingRQSpline":
        m = MaskedCouplingRQSpline(self._n_features, self.n_layers, self.hidden_size, self.num_bins, None,
                                   self.spline_range, device=self.params.device)
        m.params.copy_(self.params)
        m.tc_terms = self.tc_terms
        return m
