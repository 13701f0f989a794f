"""Optimizer resource (reference: src/flowMC/resource/optimizer.py:6-35).

The reference builds ``optax.chain(clip_by_global_norm(1.0), adamw(learning_rate, b1=momentum))``
and its state from the model's arrays.  Here ``optim`` is the same transformation as a small
config object and ``optim_state`` holds the Adam moments as flat device vectors matching the
model's flat parameter blob; the update itself is the fused ``flowmc_clip_adamw`` kernel.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from .base import Resource


@dataclass(frozen=True)
class ClipAdamW:
    """optax.chain(clip_by_global_norm(max_norm), adamw(learning_rate, b1, b2, eps, weight_decay))."""
    learning_rate: float = 1e-3
    b1: float = 0.9
    b2: float = 0.999
    eps: float = 1e-8
    weight_decay: float = 1e-4
    max_norm: float = 1.0


class OptState:
    """Adam moments (flat, same layout as the model blob) + step count."""

    def __init__(self, n_params: int, device):
        self.mu = torch.zeros(n_params, dtype=torch.float32, device=device)
        self.nu = torch.zeros(n_params, dtype=torch.float32, device=device)
        self.count = 0

    def clone(self) -> "OptState":
        s = OptState.__new__(OptState)
        s.mu, s.nu, s.count = self.mu.clone(), self.nu.clone(), self.count
        return s

    def copy_(self, other: "OptState"):
        self.mu.copy_(other.mu)
        self.nu.copy_(other.nu)
        self.count = other.count


class Optimizer(Resource):
    optim: ClipAdamW
    optim_state: OptState

    def __repr__(self):
        return "Optimizer"

    def __init__(self, model, learning_rate: float = 1e-3, momentum: float = 0.9):
        self.optim = ClipAdamW(learning_rate=learning_rate, b1=momentum)
        self.optim_state = OptState(model.params.numel(), model.params.device)

    def __call__(self, params, grads):
        raise NotImplementedError

    def print_parameters(self):
        raise NotImplementedError

    def save_resource(self, path: str):
        raise NotImplementedError

    def load_resource(self, path: str):
        raise NotImplementedError
