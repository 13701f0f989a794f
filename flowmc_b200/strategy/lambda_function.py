"""Lambda strategy (reference: src/flowMC/strategy/lambda_function.py:7-37)."""
from typing import Callable

from .base import Strategy


class Lambda(Strategy):
    """Applies a function to the resources; returns its inputs unchanged."""

    def __init__(self, lambda_function: Callable):
        self.lambda_function = lambda_function

    def __call__(self, rng_key, resources, initial_position, data):
        self.lambda_function(rng_key, resources, initial_position, data)
        return rng_key, resources, initial_position
