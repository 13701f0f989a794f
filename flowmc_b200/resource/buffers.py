"""Buffer -- device-resident sample store (reference: src/flowMC/resource/buffers.py:11-64).

Same constructor, attributes and methods as the reference.  ``data`` is a CUDA float32 tensor
initialised to -inf (buffers.py:27).  The reference's ``update_buffer`` is functional
(``dynamic_update_slice_in_dim`` copies the whole array, buffers.py:39-41); here it writes in
place, and the sampling kernels bypass it entirely by storing straight into ``data`` at the
strategy's cursor.  The silent start-index clamping of ``dynamic_update_slice`` is kept.
"""
from __future__ import annotations

import numpy as np
import torch

from .base import Resource


def _default_device():
    if not torch.cuda.is_available():
        raise RuntimeError("flowmc_b200 needs a CUDA device (there is no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def clamp_start(start: int, n_updates: int, size: int) -> int:
    """dynamic_update_slice semantics: the start index is clamped so the update fits."""
    if n_updates > size:
        raise ValueError(f"update of length {n_updates} does not fit a buffer of length {size}")
    return max(0, min(int(start), size - n_updates))


class Buffer(Resource):
    name: str
    cursor: int = 0
    cursor_dim: int = 0

    def __repr__(self):
        return "Buffer " + self.name + " with shape " + str(tuple(self.data.shape))

    @property
    def shape(self):
        return tuple(self.data.shape)

    def __init__(self, name: str, shape: tuple[int, ...], cursor_dim: int = 0, device=None):
        self.cursor_dim = cursor_dim
        self.cursor = 0
        self.name = name
        dev = _default_device() if device is None else torch.device(device)
        self.data = torch.full(tuple(shape), float("-inf"), dtype=torch.float32, device=dev)

    def __call__(self):
        return self.data

    def update_buffer(self, updates, start: int = 0):
        updates = torch.as_tensor(updates, dtype=torch.float32, device=self.data.device)
        n = updates.shape[self.cursor_dim]
        s = clamp_start(start, n, self.data.shape[self.cursor_dim])
        self.data.narrow(self.cursor_dim, s, n).copy_(updates)

    def print_parameters(self):
        print(
            f"Buffer: {self.name} with shape {tuple(self.data.shape)} and cursor"
            f" {self.cursor} at dimension {self.cursor_dim}"
        )

    def get_distribution(self, n_bins: int = 100):
        return np.histogram(self.data.detach().cpu().numpy().flatten(), bins=n_bins)

    def save_resource(self, path: str):
        np.savez(path + self.name, name=self.name, data=self.data.detach().cpu().numpy())

    def load_resource(self, path: str) -> "Buffer":
        blob = np.load(path)
        arr = blob["data"]
        result = Buffer(str(blob["name"]), arr.shape, self.cursor_dim, device=self.data.device)
        result.data.copy_(torch.from_numpy(arr))
        return result
