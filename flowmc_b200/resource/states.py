"""State -- named dictionary of str/int/bool flags (reference: src/flowMC/resource/states.py:8-63).

Strategies find their *current* target buffers through a State (take_steps.py:74-98), which is how
the bundle flips from training to production buffers.  Pure host logic.
"""
from __future__ import annotations

import numpy as np

from .base import Resource


class State(Resource):
    name: str
    data: dict

    def __repr__(self):
        return "State " + self.name + " with shape " + str(len(self.data))

    def __init__(self, data: dict, name: str = "State"):
        self.name = name
        self.data = data

    def update(self, key: list, value: list):
        for k, v in zip(key, value):
            self.data[k] = v
            print(f"Updated state {k} to {v}")

    def print_parameters(self):
        print(f"State: {self.name} with shape {len(self.data)} and data {self.data}")

    def save_resource(self, path: str):
        np.savez(path + self.name, name=self.name, data=self.data)  # type: ignore[arg-type]

    def load_resource(self, path: str) -> "State":
        blob = np.load(path, allow_pickle=True)
        return State(blob["data"].item(), str(blob["name"]))
