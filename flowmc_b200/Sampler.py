"""Sampler -- the top-level driver (reference: src/flowMC/Sampler.py:10-118).

Host orchestration only: the sampler owns a dict of resources and a dict of strategies and runs the strategies in
``strategy_order``, handing ``(rng_key, resources, position)`` from one to the next.  Everything that costs time
happens inside the strategies (one C-ABI call each on the B200 path).
"""
from __future__ import annotations

import torch

from .tracing import nvtx_range

_SETTINGS = {"verbose": False, "logging": True, "outdir": "./outdir/"}   # keyword overrides the reference accepts


class Sampler:
    def __init__(self, n_dim: int, n_chains: int, rng_key, resources: dict | None = None,
                 strategies: dict | None = None, strategy_order: list | None = None,
                 resource_strategy_bundles=None, **kwargs):
        self.n_dim, self.n_chains, self.rng_key = n_dim, n_chains, rng_key
        for name, default in _SETTINGS.items():
            setattr(self, name, kwargs.get(name, default))

        explicit = resources is not None and strategies is not None
        if explicit:
            print("Resources and strategies provided. Ignoring resource strategy bundles.")
            source = (resources, strategies, strategy_order)
        else:
            print("Resources or strategies not provided. Using resource strategy bundles.")
            if resource_strategy_bundles is None:
                raise ValueError("Resource strategy bundles not provided."
                                 "Please provide either resources and strategies or resource strategy bundles.")
            b = resource_strategy_bundles
            source = (b.resources, b.strategies, b.strategy_order)
        self.resources, self.strategies, self.strategy_order = source

    def _strategy(self, name: str):
        try:
            return self.strategies[name]
        except KeyError:
            raise ValueError(f"Invalid strategy name '{name}' provided. "
                             f"Available strategies are: {list(self.strategies.keys())}.") from None

    def sample(self, initial_position, data: dict):
        """Run every strategy of ``strategy_order`` once, in order (Sampler.py:84-108).  ``initial_position`` is
        ``[n_chains, n_dim]`` (a single position is promoted to one chain); the last strategy's position is kept in
        ``last_step``, the advanced key in ``rng_key``."""
        if not isinstance(self.strategy_order, list):
            raise AssertionError("strategy_order must be a list of strategy names")
        position = torch.atleast_2d(torch.as_tensor(initial_position, dtype=torch.float32))
        key = self.rng_key
        for name in self.strategy_order:
            with nvtx_range(f"flowmc/{name}"):       # one NVTX range per strategy call
                key, self.resources, position = self._strategy(name)(key, self.resources, position, data)
        self.rng_key, self.last_step = key, position

    def serialize(self):
        raise NotImplementedError

    def deserialize(self):
        raise NotImplementedError
