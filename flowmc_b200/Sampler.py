"""Sampler -- top-level API (reference: src/flowMC/Sampler.py:10-118).

Pure host orchestration: iterates ``strategy_order`` and threads ``(rng_key, resources,
last_step)`` through the strategies exactly like the reference (Sampler.py:84-108).
"""
from __future__ import annotations

from typing import Optional

import torch

from .resource.base import Resource
from .strategy.base import Strategy


class Sampler:
    # Essential parameters
    n_dim: int
    n_chains: int
    resources: dict
    strategies: dict
    strategy_order: Optional[list]

    # Logging hyperparameters
    verbose: bool = False
    logging: bool = True
    outdir: str = "./outdir/"

    def __init__(self, n_dim: int, n_chains: int, rng_key, resources=None, strategies=None,
                 strategy_order=None, resource_strategy_bundles=None, **kwargs):
        self.n_dim = n_dim
        self.n_chains = n_chains
        self.rng_key = rng_key

        if resources is not None and strategies is not None:
            print("Resources and strategies provided. Ignoring resource strategy bundles.")
            self.resources = resources
            self.strategies = strategies
            self.strategy_order = strategy_order
        else:
            print("Resources or strategies not provided. Using resource strategy bundles.")
            if resource_strategy_bundles is None:
                raise ValueError(
                    "Resource strategy bundles not provided."
                    "Please provide either resources and strategies or resource strategy bundles."
                )
            self.resources = resource_strategy_bundles.resources
            self.strategies = resource_strategy_bundles.strategies
            self.strategy_order = resource_strategy_bundles.strategy_order

        class_keys = list(self.__class__.__dict__.keys())
        for key, value in kwargs.items():
            if key in class_keys and not key.startswith("__"):
                setattr(self, key, value)

    def sample(self, initial_position, data: dict):
        initial_position = torch.atleast_2d(torch.as_tensor(initial_position, dtype=torch.float32))
        rng_key = self.rng_key
        last_step = initial_position
        assert isinstance(self.strategy_order, list)
        for strategy in self.strategy_order:
            if strategy not in self.strategies:
                raise ValueError(
                    f"Invalid strategy name '{strategy}' provided. "
                    f"Available strategies are: {list(self.strategies.keys())}."
                )
            rng_key, self.resources, last_step = self.strategies[strategy](
                rng_key, self.resources, last_step, data
            )
        self.rng_key = rng_key
        self.last_step = last_step

    def serialize(self):
        raise NotImplementedError

    def deserialize(self):
        raise NotImplementedError
