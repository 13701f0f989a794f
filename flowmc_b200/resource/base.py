"""Resource -- what a Strategy looks up by name in the sampler's ``resources`` dict (contract of
src/flowMC/resource/base.py:6-38).

On the B200 path a resource is one of: a ``Buffer`` (device tensor the kernels write into), a ``State`` (names of the
buffers currently targeted), a ``LogPDF`` (registered device target), a local / global kernel (parameters of a CUDA
kernel launch), a flow model (one flat device parameter blob) or an ``Optimizer`` (Adam moments on the device).
Whatever it is, it can describe its tunable parameters and, where that makes sense, persist itself.
"""
from __future__ import annotations

import abc


class Resource(abc.ABC):
    def _unsupported(self, what: str):
        return NotImplementedError(f"{type(self).__name__} does not implement {what}")

    @abc.abstractmethod
    def print_parameters(self) -> None:
        """Print the tunable parameters (the reference's tests assert some of these strings)."""

    @abc.abstractmethod
    def save_resource(self, path: str) -> None:
        """Write the resource under ``path`` (a prefix: implementations append their own name / suffix)."""

    @abc.abstractmethod
    def load_resource(self, path: str):
        """Return the resource stored under ``path``."""
