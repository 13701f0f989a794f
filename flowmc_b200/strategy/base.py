"""Strategy ABC (reference: src/flowMC/strategy/base.py:8-31)."""
from abc import ABC, abstractmethod


class Strategy(ABC):
    """A callable block ``(rng_key, resources, initial_position, data) ->
    (rng_key, resources, position)`` that the Sampler runs in ``strategy_order``."""

    @abstractmethod
    def __init__(self):
        raise NotImplementedError

    @abstractmethod
    def __call__(self, rng_key, resources, initial_position, data):
        raise NotImplementedError
