"""AdamOptimization (reference: src/flowMC/strategy/optimization.py:12-164).

Same constructor, attributes, ``repr``, call contract and error messages.  Where the reference runs
``vmap(scan(grad -> optax.adam -> projection_box))`` (optimization.py:118-153), this makes ONE C-ABI call
(``flowmc_adam_optimize``): a CUDA kernel runs all ``n_steps`` for every chain with the chain's position, both Adam
moments and its key in registers, the target's analytic gradient and jax.random-compatible keys.

``logpdf`` is a ``LogPDF`` resource or a ``DeviceTarget`` (a registered device function): a Python callable cannot run
in the kernel (``TypeError``; there is no CPU path).  ``optimize(rng_key, objective, ...)`` keeps the reference's
signature, but the objective is by construction ``-logpdf``: pass ``None`` (or anything) -- it is ignored.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .._lib import check, lib
from ..resource.logPDF import LogPDF
from ..targets import DeviceTarget
from .base import Strategy

_u32p = C.POINTER(C.c_uint32)


def adam_bias_corrections(n_steps: int, b1: float = 0.9, b2: float = 0.999) -> np.ndarray:
    """float32 [n_steps, 2]: 1 - b^t for t = 1..n_steps, evaluated like optax's ``1 - decay ** count`` in float32."""
    t = np.arange(1, n_steps + 1, dtype=np.float32)
    return np.stack([np.float32(1) - np.float32(b1) ** t, np.float32(1) - np.float32(b2) ** t], axis=1).astype(np.float32)


class AdamOptimization(Strategy):
    """Optimize a set of chains using Adam optimization (see the reference docstring, optimization.py:13-29).

    Args:
        logpdf: LogPDF resource / DeviceTarget to maximise.
        n_steps: number of optimization steps.
        learning_rate: Adam learning rate.
        noise_level: the gradient is multiplied by ``1 + normal() * noise_level`` at every step.
        bounds: ``(n_dim, 2)`` or ``(1, 2)`` (broadcast) box the positions are projected to after every step.
    """

    def __repr__(self):
        return "AdamOptimization"

    def __init__(self, logpdf, n_steps: int = 100, learning_rate: float = 1e-2, noise_level: float = 10,
                 bounds=np.array([[-np.inf, np.inf]])):
        if isinstance(logpdf, DeviceTarget):
            pass
        elif isinstance(logpdf, LogPDF):
            pass
        else:
            raise TypeError("AdamOptimization needs a LogPDF resource or a DeviceTarget (a registered device function "
                            "with an analytic gradient); Python callables cannot run in the CUDA kernel")
        self.logpdf = logpdf
        self.n_steps = n_steps
        self.learning_rate = learning_rate
        self.noise_level = noise_level
        bounds = np.asarray(bounds.detach().cpu() if isinstance(bounds, torch.Tensor) else bounds, dtype=np.float32)
        self.bounds = bounds
        if bounds.ndim != 2 or bounds.shape[1] != 2:
            raise ValueError(f"bounds must have shape (n_dim, 2) or (1, 2), got {bounds.shape}")
        # global chain shard owned by this process: (offset, n_chains_global) or None = all chains
        self.chain_shard = None

    def set_chain_shard(self, offset: int, n_chains_global: int):
        self.chain_shard = (int(offset), int(n_chains_global))

    @property
    def _target(self) -> DeviceTarget:
        return self.logpdf.target if isinstance(self.logpdf, LogPDF) else self.logpdf

    def __call__(self, rng_key, resources, initial_position, data):
        rng_key, optimized_positions, _ = self.optimize(rng_key, None, initial_position, data)
        return rng_key, resources, optimized_positions

    def optimize(self, rng_key, objective, initial_position, data):
        """Returns (rng_key, optimized_positions [n_chain, n_dim], final_log_prob [n_chain])."""
        x0 = torch.as_tensor(initial_position, dtype=torch.float32)
        if not x0.is_cuda:
            if not torch.cuda.is_available():
                raise RuntimeError("flowmc_b200 needs a CUDA device (there is no CPU fallback)")
            x0 = x0.cuda()
        x0 = x0.contiguous()
        n, n_dim = x0.shape
        if not (self.bounds.shape[0] == 1 or self.bounds.shape[0] == n_dim):
            raise ValueError(
                f"bounds shape {self.bounds.shape} is incompatible with n_dim={n_dim}. "
                "Provide bounds of shape (1, 2) for broadcasting or (n_dim, 2) for per-dimension bounds.")
        print("Using Adam optimization")
        dev = x0.device
        b = np.broadcast_to(self.bounds, (n_dim, 2))
        lo = torch.from_numpy(np.ascontiguousarray(b[:, 0])).to(dev)
        hi = torch.from_numpy(np.ascontiguousarray(b[:, 1])).to(dev)
        bc = torch.from_numpy(adam_bias_corrections(int(self.n_steps))).to(dev)
        tgt = self._target
        pk = tgt.packed_on(data, n_dim, dev)
        key = np.ascontiguousarray(rng_key, dtype=np.uint32)
        key_out = np.zeros(2, np.uint32)
        out = torch.empty_like(x0)
        lp = torch.empty(n, dtype=torch.float32, device=dev)
        offset, n_glob = self.chain_shard if self.chain_shard is not None else (0, n)
        with torch.cuda.device(dev):
            check(lib.flowmc_adam_optimize(tgt.target_id, pk.data_ptr(), key.ctypes.data_as(_u32p), x0.data_ptr(), n,
                                           n_dim, int(self.n_steps), float(self.learning_rate), float(self.noise_level),
                                           lo.data_ptr(), hi.data_ptr(), bc.data_ptr(), offset, n_glob,
                                           key_out.ctypes.data_as(_u32p), out.data_ptr(), lp.data_ptr(),
                                           torch.cuda.current_stream().cuda_stream))
        if bool(torch.isinf(lp).any()) or bool(torch.isnan(lp).any()):
            print("Warning: Optimization accessed infinite or NaN log-probabilities.")
        return key_out, out, lp
