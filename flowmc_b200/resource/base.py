"""Resource ABC -- same contract as src/flowMC/resource/base.py:6-38."""
from abc import ABC, abstractmethod


class Resource(ABC):
    """Objects a Strategy looks up by name: kernels, models, buffers, states, optimisers."""

    @abstractmethod
    def __init__(self):
        raise NotImplementedError

    @abstractmethod
    def print_parameters(self):
        raise NotImplementedError

    @abstractmethod
    def save_resource(self, path: str):
        raise NotImplementedError

    @abstractmethod
    def load_resource(self, path: str):
        raise NotImplementedError
