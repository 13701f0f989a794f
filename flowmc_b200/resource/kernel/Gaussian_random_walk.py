"""Gaussian random walk local kernel (reference: src/flowMC/resource/kernel/Gaussian_random_walk.py:10-71)."""
from __future__ import annotations

from ..._lib import LocalParams
from .base import LocalKernel


class GaussianRandomWalk(LocalKernel):
    """Gaussian random walk sampler class."""

    KIND = 2

    def __repr__(self):
        return "Gaussian Random Walk with step size " + str(self.step_size)

    def __init__(self, step_size: float):
        super().__init__()
        self.step_size = step_size

    def _local_params(self, n_dims, device):
        p = LocalParams()
        p.step_size = float(self.step_size)
        p.layout_hint = int(self.layout_hint)
        p.force_n_seg = int(self.force_n_seg)
        p.slots_override = int(self.slots_override)
        return p, []

    def print_parameters(self):
        print("Gaussian Random Walk parameters:")
        print(f"step_size: {self.step_size}")

    def save_resource(self, path):
        raise NotImplementedError

    def load_resource(self, path):
        raise NotImplementedError
