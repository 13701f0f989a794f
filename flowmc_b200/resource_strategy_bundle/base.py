"""ResourceStrategyBundle (reference: src/flowMC/resource_strategy_bundle/base.py:7-17): three
attributes -- ``resources``, ``strategies``, ``strategy_order`` -- that a Sampler consumes."""
from abc import ABC


class ResourceStrategyBundle(ABC):
    resources: dict
    strategies: dict
    strategy_order: list
