"""Device targets: the B200 replacement for the reference's Python ``logpdf(x, data)`` callables.

The reference differentiates an arbitrary Python function with ``jax.value_and_grad``
(src/flowMC/resource/logPDF.py:60-61, resource/kernel/MALA.py:59).  On the B200 path a target is
a *registered device function* with an analytic gradient (``include/flowmc_target.cuh``); a
``DeviceTarget`` names one and knows how to pack the reference's ``data`` dict into the flat
float32 parameter block the device function reads.  Python callables cannot run in the
kernels -- ``LogPDF`` raises if it is given one (no CPU fallback by design).
"""
from __future__ import annotations

import os
import tempfile
from pathlib import Path
from typing import Callable, Optional

import numpy as np
import torch

from ._lib import check, lib, load_plugin


def _as_np(v) -> np.ndarray:
    if isinstance(v, torch.Tensor):
        v = v.detach().cpu().numpy()
    return np.asarray(v, dtype=np.float32)


class DeviceTarget:
    """A named, registered device log-density plus the host-side packing of its parameters."""

    def __init__(self, name: str, pack: Callable[[Optional[dict], int], np.ndarray], description: str = ""):
        self.name = name
        self.target_id = check(lib.flowmc_target_lookup(name.encode()))
        self._pack = pack
        self.description = description or name
        self._cache: dict = {}

    def __repr__(self):
        return f"DeviceTarget({self.description})"

    def pack(self, data: Optional[dict], n_dims: int) -> np.ndarray:
        return np.ascontiguousarray(self._pack(data, n_dims), dtype=np.float32).reshape(-1)

    def packed_on(self, data: Optional[dict], n_dims: int, device: torch.device) -> torch.Tensor:
        """Packed parameter block on ``device`` (cached per data object and device)."""
        k = (id(data), n_dims, str(device))
        hit = self._cache.get(k)
        if hit is not None and hit[0] is data:
            return hit[1]
        t = torch.from_numpy(self.pack(data, n_dims)).to(device)
        self._cache = {k: (data, t)}
        return t

    # value / value_and_grad on the GPU -----------------------------------------------------
    def evaluate(self, x: torch.Tensor, data: Optional[dict], want_grad: bool = False):
        if not (isinstance(x, torch.Tensor) and x.is_cuda):
            raise TypeError("DeviceTarget.evaluate needs a CUDA float32 tensor (there is no CPU path)")
        single = x.dim() == 1
        x2 = x.reshape(1, -1) if single else x
        x2 = x2.contiguous().float()
        n, d = x2.shape
        pk = self.packed_on(data, d, x2.device)
        lp = torch.empty(n, dtype=torch.float32, device=x2.device)
        g = torch.empty_like(x2) if want_grad else None
        with torch.cuda.device(x2.device):
            check(lib.flowmc_target_eval(self.target_id, pk.data_ptr(), x2.data_ptr(), n, d, lp.data_ptr(),
                                         g.data_ptr() if want_grad else None,
                                         torch.cuda.current_stream().cuda_stream))
        if single:
            lp = lp[0]
            g = g[0] if want_grad else None
        return (lp, g) if want_grad else lp

    def __call__(self, x, data=None):
        return self.evaluate(x, data, want_grad=False)


def _get(data, key, default):
    if data is None or key is None:
        return default
    v = data.get(key, None)
    return default if v is None else v


def iso_gaussian(c: float = 0.5, data_key: Optional[str] = "data") -> DeviceTarget:
    """logp = -c * |x - data[data_key]|^2  (mean 0 if the key is absent).  c=0.5: the reference's
    test targets (test/unit/test_kernels.py:14-15, test/integration/test_quickstart.py:7-8)."""
    def pack(data, d):
        mu = _as_np(_get(data, data_key, np.zeros(d))).reshape(-1)
        return np.concatenate([[np.float32(c)], mu.astype(np.float32)])
    return DeviceTarget("iso_gaussian", pack, f"iso_gaussian(c={c})")


def dual_moon(data_key: Optional[str] = None) -> DeviceTarget:
    """docs/tutorials/dualmoon.ipynb:77-84 (data_key=None) / test/integration/test_MALA.py:14-23 ("data")."""
    def pack(data, d):
        return _as_np(_get(data, data_key, np.zeros(d))).reshape(-1)
    return DeviceTarget("dual_moon", pack, "dual_moon")


def ar1_gaussian(rho: float = 0.9) -> DeviceTarget:
    return DeviceTarget("ar1_gaussian", lambda data, d: np.array([rho], np.float32), f"ar1_gaussian(rho={rho})")


def dense_gaussian(precision) -> DeviceTarget:
    P = _as_np(precision)
    return DeviceTarget("dense_gaussian", lambda data, d: P.reshape(-1), "dense_gaussian")


def rosenbrock() -> DeviceTarget:
    return DeviceTarget("rosenbrock", lambda data, d: np.zeros(1, np.float32), "rosenbrock")


def gaussian_mixture(means, inv_var: float = 1.0, logw=None) -> DeviceTarget:
    mu = _as_np(means)
    K = mu.shape[0]
    if K > 8:
        raise ValueError("gaussian_mixture supports at most 8 components")
    lw = np.full(K, -np.log(K), np.float32) if logw is None else _as_np(logw)
    blk = np.concatenate([[np.float32(K), np.float32(inv_var)], lw, mu.reshape(-1)]).astype(np.float32)
    return DeviceTarget("gaussian_mixture", lambda data, d: blk, f"gaussian_mixture(K={K})")


def compile_target(source: str, name: str, pack: Callable[[Optional[dict], int], np.ndarray],
                   build_dir: Optional[str] = None) -> DeviceTarget:
    """Compile a user plugin (CUDA source text that includes ``flowmc_target.cuh`` and calls
    FLOWMC_REGISTER_TARGET(..., "<name>")), load it and return its DeviceTarget."""
    from .build import build_plugin
    bdir = Path(build_dir or tempfile.mkdtemp(prefix="flowmc_plugin_"))
    bdir.mkdir(parents=True, exist_ok=True)
    src = bdir / f"{name}.cu"
    src.write_text(source)
    so = build_plugin(src, bdir / f"libflowmc_target_{name}.so")
    load_plugin(os.fspath(so))
    return DeviceTarget(name, pack, name)
