"""HMC local kernel (reference: src/flowMC/resource/kernel/HMC.py:11-163).

``condition_matrix`` is the inverse mass matrix and must be a 2-D matrix, exactly as in the
reference (HMC.py:135 calls ``jnp.linalg.inv`` on it; the default scalar ``1`` cannot run there
either).  The host precomputes the two things the kernel needs from it: ``L = chol(inv(M))``
(momentum draw, HMC.py:133-136) and the column sums of ``M`` (kinetic energy
``0.5 * (p**2 * M).sum()`` and its gradient, HMC.py:121).
"""
from __future__ import annotations

import numpy as np
import torch

from ..._lib import LocalParams
from .base import LocalKernel


class HMC(LocalKernel):
    """Hamiltonian Monte Carlo sampler class."""

    KIND = 1

    @property
    def n_leapfrog(self) -> int:
        return self.leapfrog_coefs.shape[0] - 2

    def __repr__(self):
        return "HMC with step size " + str(self.step_size) + " and " + str(self.n_leapfrog) + " leapfrog steps"

    def __init__(self, condition_matrix=1, step_size: float = 0.1, n_leapfrog: int = 10):
        super().__init__()
        self.condition_matrix = condition_matrix
        self.step_size = step_size
        coefs = np.ones((n_leapfrog + 2, 2), dtype=np.float32)
        coefs[0] = (0.0, 0.5)
        coefs[-1] = (1.0, 0.5)
        self.leapfrog_coefs = coefs
        self._dev_cache = None

    def _host_constants(self):
        M = self.condition_matrix
        if isinstance(M, torch.Tensor):
            M = M.detach().cpu().numpy()
        M = np.asarray(M, dtype=np.float32)
        if M.ndim != 2 or M.shape[0] != M.shape[1]:
            raise ValueError("HMC condition_matrix must be a square 2-D matrix (HMC.py:135 inverts it)")
        L = np.linalg.cholesky(np.linalg.inv(M.astype(np.float64))).astype(np.float32)
        colsum = M.sum(axis=0, dtype=np.float32).astype(np.float32)
        diag = bool(np.count_nonzero(L - np.diag(np.diag(L))) == 0)
        return L, colsum, diag

    def _local_params(self, n_dims, device):
        if self._dev_cache is None or self._dev_cache[0] != str(device):
            L, colsum, diag = self._host_constants()
            if L.shape[0] != n_dims:
                raise ValueError(f"condition_matrix is {L.shape}, chains have {n_dims} dimensions")
            self._dev_cache = (str(device), torch.from_numpy(L).contiguous().to(device),
                               torch.from_numpy(colsum).to(device), diag)
        _, Ld, cs, diag = self._dev_cache
        p = LocalParams()
        p.step_size = float(self.step_size)
        p.n_leapfrog = int(self.n_leapfrog)
        p.hmc_chol = Ld.data_ptr()
        p.hmc_colsum = cs.data_ptr()
        p.hmc_chol_diagonal = 1 if diag else 0
        p.layout_hint = int(self.layout_hint)
        p.force_n_seg = int(self.force_n_seg)
        p.slots_override = int(self.slots_override)
        return p, [Ld, cs]

    def print_parameters(self):
        print("HMC parameters:")
        print(f"step_size: {self.step_size}")
        print(f"n_leapfrog: {self.n_leapfrog}")
        print(f"condition_matrix: {self.condition_matrix}")

    def save_resource(self, path):
        raise NotImplementedError

    def load_resource(self, path):
        raise NotImplementedError
