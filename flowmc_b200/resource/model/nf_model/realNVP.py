"""RealNVP and AffineCoupling (reference: src/flowMC/resource/model/nf_model/realNVP.py:18-228).

Same constructors and methods as the reference.  Like the spline flow, the model is ONE flat float32 device vector
(``self.params``, layout ``FlowmcRealNVPDesc`` in include/flowmc_b200.h) so that the fused clip + AdamW and the
data-parallel gradient all-reduce each see a single buffer; ``forward`` / ``inverse`` / ``log_prob`` / ``sample`` /
the training loss gradient are one C-ABI call each into csrc/flow_realnvp.cu.  Methods take batches ``[n, d]`` (the
reference's per-sample methods are always used under ``vmap``); a single sample ``[d]`` returns un-batched results.

Initialisation reproduces the reference's key schedule bit for bit (realNVP.py:157-170, common.py:83-107,
equinox.nn.Linear's uniform(+-1/sqrt(in)) init).  The coupling masks are part of the parameter blob because they are
float leaves of the reference model: they receive zero gradient but AdamW's weight decay.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from .... import random as frandom
from ...._lib import RealNVPDesc, check, lib
from .base import NFModel

_u32p = C.POINTER(C.c_uint32)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _mlp_init(model, layer: int, which: int, key: np.ndarray, scale: float = 1e-4):
    """MLP([d, h, d], key, scale) (common.py:83-107) into the blob views of coupling layer ``layer``."""
    d, h = model._n_features, model.n_hidden
    dev = model.params.device
    k, sub1, sub2 = frandom.split(key, 3)
    _, bkey = frandom.split(sub1, 2)                                   # eqx.nn.Linear: wkey, bkey = split(key, 2)
    lim = 1.0 / math.sqrt(d)
    model.b1(layer, which).copy_(frandom.uniform(bkey, (h,), -lim, lim, device=dev))
    std = float(np.sqrt(np.float32(scale / d)))                         # jnp.sqrt(scale / shape[i]) in float32
    model.W1(layer, which).copy_(frandom.normal(sub2, (h, d), device=dev) * std)
    k, sub = frandom.split(k, 2)
    wkey, bkey = frandom.split(sub, 2)
    lim = 1.0 / math.sqrt(h)
    model.W2(layer, which).copy_(frandom.uniform(wkey, (d, h), -lim, lim, device=dev))
    model.b2(layer, which).copy_(frandom.uniform(bkey, (d,), -lim, lim, device=dev))


class _NVPBlob:
    """Views into the flat parameter vector shared by RealNVP and AffineCoupling."""

    def _alloc(self, n_features: int, n_layers: int, n_hidden: int, dt: float, device):
        if not torch.cuda.is_available():
            raise RuntimeError("flowmc_b200 needs a CUDA device (there is no CPU fallback)")
        self._n_features, self.n_layers, self.n_hidden, self.dt = int(n_features), int(n_layers), int(n_hidden), float(dt)
        self.desc = RealNVPDesc()
        check(lib.flowmc_realnvp_desc_init(C.byref(self.desc), self._n_features, self.n_layers, self.n_hidden, self.dt))
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.params = torch.zeros(int(self.desc.n_params), dtype=torch.float32, device=dev)
        d = self._n_features
        self._view(self.desc.off_data_cov, (d, d)).copy_(torch.eye(d, device=dev))
        self._view(self.desc.off_base_cov, (d, d)).copy_(torch.eye(d, device=dev))

    def _view(self, off: int, shape, layer: int | None = None) -> torch.Tensor:
        o = int(off) + (0 if layer is None else layer * int(self.desc.layer_stride))
        return self.params[o:o + math.prod(shape)].view(*shape)

    def W1(self, layer: int, which: int) -> torch.Tensor:
        """First Linear of the scale (which = 0) / shift (1) MLP of coupling layer ``layer``: [h, d]."""
        return self._view(self.desc.off_W1t if which else self.desc.off_W1s, (self.n_hidden, self._n_features), layer)

    def b1(self, layer: int, which: int) -> torch.Tensor:
        return self._view(self.desc.off_b1t if which else self.desc.off_b1s, (self.n_hidden,), layer)

    def W2(self, layer: int, which: int) -> torch.Tensor:
        return self._view(self.desc.off_W2t if which else self.desc.off_W2s, (self._n_features, self.n_hidden), layer)

    def b2(self, layer: int, which: int) -> torch.Tensor:
        return self._view(self.desc.off_b2t if which else self.desc.off_b2s, (self._n_features,), layer)

    def layer_mask(self, layer: int) -> torch.Tensor:
        """MaskedCouplingLayer._mask of coupling layer ``layer`` (float: 1 = unchanged / conditioning)."""
        return self._view(self.desc.off_mask, (self._n_features,), layer)

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
        n = x2.shape[0]
        y = torch.empty_like(x2)
        ld = torch.empty(n, dtype=torch.float32, device=x2.device)
        with torch.cuda.device(x2.device):
            check(fn(C.byref(self.desc), self.params.data_ptr(), x2.data_ptr(), n, y.data_ptr(), ld.data_ptr(),
                     _stream()))
        if single:
            return y[0], ld[0]
        return y.reshape(*lead, x2.shape[-1]), ld.reshape(*lead)


class AffineCoupling(_NVPBlob):
    """Affine coupling layer (realNVP.py:18-100): the masked entries (mask = 1) are transformed, conditioned on
    ``x * (1 - mask)``; ``s = tanh(mask * scale_MLP(.)) * dt``, ``t = mask * translate_MLP(.) * dt``,
    ``y = (x + t) * exp(s)``, ``log_det = sum(s)``.  Runs as a one-layer blob of the RealNVP kernels with the
    complementary coupling mask (for a 0 / 1 mask the two formulations are the same function)."""

    def __init__(self, n_features: int, n_hidden: int, mask, key, dt: float = 1, scale: float = 1e-4, device=None):
        self._alloc(n_features, 1, n_hidden, dt, device)
        m = torch.as_tensor(np.asarray(mask, dtype=np.float32))
        if not bool(((m == 0) | (m == 1)).all()):
            raise NotImplementedError("flowmc_b200 AffineCoupling takes a 0 / 1 mask")
        self._mask = m.to(self.params.device)
        self.layer_mask(0).copy_(1.0 - self._mask)
        if key is not None:
            _, scale_subkey, translate_subkey = frandom.split(np.asarray(key, dtype=np.uint32), 3)   # realNVP.py:46
            _mlp_init(self, 0, 0, scale_subkey, scale)
            _mlp_init(self, 0, 1, translate_subkey, scale)

    @property
    def mask(self):
        return self._mask

    @property
    def n_features(self):
        return self._n_features

    def __call__(self, x):
        return self.forward(x)

    def forward(self, x):
        return self._transform(lib.flowmc_realnvp_forward, x)

    def inverse(self, x):
        return self._transform(lib.flowmc_realnvp_inverse, x)


class RealNVP(_NVPBlob, NFModel):
    """RealNVP flow (realNVP.py:102-228): ``n_layers`` masked affine couplings with relu MLP conditioners."""

    def __repr__(self):
        return "RealNVP with n_features=" + str(self._n_features) + ", n_layers=" + str(self.n_layers)

    def __init__(self, n_features: int, n_layers: int, n_hidden: int, key, device=None, **kwargs):
        if kwargs.get("base_dist") is not None:
            raise NotImplementedError("flowmc_b200 supports the default Gaussian base distribution")
        self._alloc(n_features, n_layers, n_hidden, 1.0, device)
        d = self._n_features
        if kwargs.get("data_mean") is not None:
            self.data_mean.copy_(torch.as_tensor(kwargs["data_mean"], dtype=torch.float32))
        if kwargs.get("data_cov") is not None:
            self.data_cov.copy_(torch.atleast_2d(torch.as_tensor(kwargs["data_cov"], dtype=torch.float32)))
        for l in range(self.n_layers):                                 # realNVP.py:160-163
            m = torch.ones(d)
            m[: int(d / 2)] = 0
            self.layer_mask(l).copy_(1 - m if l % 2 == 0 else m)
        if key is not None:
            keys = frandom.split(np.asarray(key, dtype=np.uint32), self.n_layers)   # realNVP.py:169
            for l in range(self.n_layers):
                _, scale_subkey, shift_subkey = frandom.split(keys[l], 3)          # realNVP.py:159
                _mlp_init(self, l, 0, scale_subkey)
                _mlp_init(self, l, 1, shift_subkey)

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

    def forward(self, x, key=None, condition=None):
        """realNVP.py:172-188: returns (y, log_det), no whitening."""
        return self._transform(lib.flowmc_realnvp_forward, x)

    def inverse(self, x, condition=None):
        """realNVP.py:190-206."""
        return self._transform(lib.flowmc_realnvp_inverse, x)

    def __call__(self, x):
        return self.forward(x)

    def log_prob(self, x):
        """realNVP.py:214-221."""
        x2, single, lead = self._prep(x)
        n = x2.shape[0]
        lp = torch.empty(n, dtype=torch.float32, device=x2.device)
        with torch.cuda.device(x2.device):
            check(lib.flowmc_realnvp_log_prob(C.byref(self.desc), self.params.data_ptr(), x2.data_ptr(), n,
                                              lp.data_ptr(), _stream()))
        return lp[0] if single else lp.reshape(*lead)

    def sample(self, rng_key, n_samples: int):
        """realNVP.py:208-212: base.sample(key, n) -> inverse -> un-whiten."""
        key = np.ascontiguousarray(rng_key, dtype=np.uint32)
        out = torch.empty((int(n_samples), self._n_features), dtype=torch.float32, device=self.params.device)
        if int(n_samples) == 0:
            return out
        with torch.cuda.device(out.device):
            check(lib.flowmc_realnvp_sample(C.byref(self.desc), self.params.data_ptr(), None,
                                            key.ctypes.data_as(_u32p), int(n_samples), int(n_samples), out.data_ptr(),
                                            _stream()))
        return out

    def sample_rows(self, keys_d: torch.Tensor, rows_per_key: int) -> torch.Tensor:
        """Row r = ``sample(keys[r // rows_per_key], rows_per_key)[r % rows_per_key]`` for device keys [n_keys, 2]
        (what NFProposal draws for n_keys chains in one launch)."""
        n = int(keys_d.shape[0]) * int(rows_per_key)
        out = torch.empty((n, self._n_features), dtype=torch.float32, device=self.params.device)
        with torch.cuda.device(out.device):
            check(lib.flowmc_realnvp_sample(C.byref(self.desc), self.params.data_ptr(), keys_d.data_ptr(), None,
                                            int(rows_per_key), n, out.data_ptr(), _stream()))
        return out

    # ---- training hooks of NFModel ---------------------------------------------------------------------------
    def _loss_grad_workspace_bytes(self, n_rows: int) -> int:
        return int(lib.flowmc_realnvp_loss_grad_workspace_bytes(C.byref(self.desc), int(n_rows)))

    def _loss_grad_call(self, x_ptr, idx_ptr, n, inv_n_total, grad_ptr, loss_ptr, ws_ptr, ws_bytes, stream):
        return lib.flowmc_realnvp_loss_grad(C.byref(self.desc), self.params.data_ptr(), x_ptr, idx_ptr, n, inv_n_total,
                                            grad_ptr, loss_ptr, ws_ptr, ws_bytes, stream)

    # ---- resource ----------------------------------------------------------------------------------------------
    def print_parameters(self):
        print("RealNVP parameters:")
        print(f"Data mean: {self.data_mean}")
        print(f"Data covariance: {self.data_cov}")

    def save_model(self, path: str):
        """Flat blob + architecture header (npz).  The reference's ``.eqx`` leaf order for RealNVP is not restated
        (only the spline flow's is, flowmc_b200/eqx_io.py)."""
        np.savez(path + ".npz", params=self.params.detach().cpu().numpy(),
                 arch=np.array([self._n_features, self.n_layers, self.n_hidden]))

    def load_model(self, path: str) -> "RealNVP":
        z = np.load(path + ".npz")
        if tuple(int(v) for v in z["arch"]) != (self._n_features, self.n_layers, self.n_hidden):
            raise ValueError("saved RealNVP has a different architecture")
        m = self.clone()
        m.params.copy_(torch.from_numpy(z["params"]).to(self.params.device))
        return m

    save_resource = save_model
    load_resource = load_model

    def clone(self) -> "RealNVP":
        m = RealNVP(self._n_features, self.n_layers, self.n_hidden, None, device=self.params.device)
        m.params.copy_(self.params)
        return m
