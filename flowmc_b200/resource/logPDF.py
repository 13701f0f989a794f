"""LogPDF / Variable (reference: src/flowMC/resource/logPDF.py:9-81).

Same constructor ``LogPDF(log_pdf, variables=None, n_dims=None)``.  ``log_pdf`` must be a
``flowmc_b200.targets.DeviceTarget`` (or the name of a registered target): the kernels call the
target as a compiled device function with an analytic gradient, so an arbitrary Python callable
cannot be used -- it raises instead of silently running on the CPU.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

from ..targets import DeviceTarget
from .base import Resource


@dataclass
class Variable:
    name: str
    continuous: bool


class LogPDF(Resource):
    def __init__(self, log_pdf, variables: Optional[list] = None, n_dims: Optional[int] = None):
        if isinstance(log_pdf, LogPDF):
            log_pdf = log_pdf.log_pdf
        if isinstance(log_pdf, str):
            log_pdf = DeviceTarget(log_pdf, lambda data, d: [0.0], log_pdf)
        if not isinstance(log_pdf, DeviceTarget):
            raise TypeError(
                "flowmc_b200.LogPDF needs a DeviceTarget (flowmc_b200.targets.*, or "
                "targets.compile_target for your own CUDA plugin): the B200 kernels call the target as a "
                "registered device function with an analytic gradient; Python callables cannot run there "
                "and there is no CPU fallback."
            )
        self.log_pdf = log_pdf
        if variables is None and n_dims is not None:
            self.variables = [Variable("x_" + str(i), True) for i in range(n_dims)]
        elif variables is not None:
            self.variables = variables
        else:
            raise ValueError("Either variables or n_dims must be provided")

    @property
    def n_dims(self):
        return len(self.variables)

    @property
    def target(self) -> DeviceTarget:
        return self.log_pdf

    def __repr__(self):
        return "LogPDF with " + str(self.n_dims) + " dimensions"

    def __call__(self, x, data=None):
        return self.log_pdf.evaluate(x, data, want_grad=False)

    def value_and_grad(self, x, data=None):
        return self.log_pdf.evaluate(x, data, want_grad=True)

    def print_parameters(self):
        print("LogPDF with variables:")
        for var in self.variables:
            print(var.name, var.continuous)

    def save_resource(self, path):
        raise NotImplementedError

    def load_resource(self, path):
        raise NotImplementedError
