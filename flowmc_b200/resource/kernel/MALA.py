"""MALA local kernel (reference: src/flowMC/resource/kernel/MALA.py:10-99)."""
from __future__ import annotations

from ..._lib import LocalParams
from .base import LocalKernel


class MALA(LocalKernel):
    """Metropolis-adjusted Langevin algorithm sampler class."""

    KIND = 0

    def __repr__(self):
        return "MALA with step size " + str(self.step_size)

    def __init__(self, step_size: float):
        super().__init__()
        if hasattr(step_size, "shape") and tuple(getattr(step_size, "shape")) not in ((), (1,)):
            raise NotImplementedError("flowmc_b200 MALA supports a scalar step_size (MALA.py:68 dt: Float)")
        self.step_size = step_size

    def _local_params(self, n_dims, device):
        p = LocalParams()
        p.step_size = float(self.step_size)
        p.layout_hint = int(self.layout_hint)
        p.force_n_seg = int(self.force_n_seg)
        p.slots_override = int(self.slots_override)
        return p, []

    def print_parameters(self):
        print("MALA parameters:")
        print(f"step_size: {self.step_size}")

    def save_resource(self, path):
        raise NotImplementedError

    def load_resource(self, path):
        raise NotImplementedError
