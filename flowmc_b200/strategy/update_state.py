"""UpdateState strategy (reference: src/flowMC/strategy/update_state.py:7-47)."""
from ..resource.states import State
from .base import Strategy


class UpdateState(Strategy):
    """Update a State resource in place (e.g. switch the target buffers to production)."""

    def __init__(self, state_name: str, keys: list, values: list):
        self.state_name = state_name
        self.keys = keys
        self.values = values

    def __call__(self, rng_key, resources, initial_position, data):
        assert isinstance(state := resources[self.state_name], State), \
            f"Resource {self.state_name} is not a State resource."
        state.update(self.keys, self.values)
        return rng_key, resources, initial_position
