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


class BoxQuadraticPrior:
    """Device-evaluable log-prior of the fixed-function family the tempered kernels support:

        log_prior(x) = -sum_j c_j (x_j - m_j)^2   if lo_j <= x_j <= hi_j for every j,   -inf otherwise.

    ``c = 1 / (2 sigma^2)`` is an (unnormalised) Gaussian prior, ``c = 0`` with finite bounds a uniform prior, the
    defaults the flat prior 0.  Scalars broadcast over the dimensions.  (The reference takes an arbitrary Python
    callable, resource/logPDF.py:88-100; on the B200 path the prior runs inside the sampling kernel.)"""

    def __init__(self, c=0.0, mean=0.0, lower=-float("inf"), upper=float("inf")):
        self.c, self.mean, self.lower, self.upper = c, mean, lower, upper

    def packed(self, n_dims: int):
        import numpy as np
        rows = [np.broadcast_to(np.asarray(v, dtype=np.float32), (n_dims,)) for v in
                (self.c, self.mean, self.lower, self.upper)]
        return np.ascontiguousarray(np.stack(rows), dtype=np.float32)      # [4, d] = c, m, lo, hi

    def is_flat(self) -> bool:
        import numpy as np
        return bool(np.all(np.asarray(self.c) == 0) and np.all(np.isneginf(np.asarray(self.lower, dtype=np.float64)))
                    and np.all(np.isposinf(np.asarray(self.upper, dtype=np.float64))))

    def __call__(self, x, data=None):
        import torch
        p = torch.from_numpy(self.packed(x.shape[-1])).to(x.device)
        r = x - p[1]
        val = -(p[0] * r * r).sum(dim=-1)
        inside = ((x >= p[2]) & (x <= p[3])).all(dim=-1)
        return torch.where(inside, val, torch.full_like(val, float("-inf")))


class TemperedPDF(LogPDF):
    """Reference: src/flowMC/resource/logPDF.py:84-106.  ``tempered_log_pdf(T, x, data) = (1 / T) * log_likelihood(x,
    data) + log_prior(x, data)``.  ``log_likelihood`` is a DeviceTarget like every LogPDF here; ``log_prior`` is a
    ``BoxQuadraticPrior`` (or None / 0 for the flat prior the reference's tests use)."""

    def __init__(self, log_likelihood, log_prior=None, variables=None, n_dims=None, n_temps=5, max_temp=100):
        super().__init__(log_likelihood, variables, n_dims)
        if log_prior is None or (isinstance(log_prior, (int, float)) and log_prior == 0):
            log_prior = BoxQuadraticPrior()
        if not isinstance(log_prior, BoxQuadraticPrior):
            raise TypeError("flowmc_b200.TemperedPDF needs a BoxQuadraticPrior (Gaussian / uniform / flat) as log_prior: "
                            "the prior is evaluated inside the CUDA kernels; Python callables cannot run there")
        self.log_prior = log_prior

    def __call__(self, x, data=None):
        return super().__call__(x, data)

    def tempered_log_pdf(self, temperatures, x, data=None):
        import torch
        base_pdf = super().__call__(x, data)
        t = torch.as_tensor(temperatures, dtype=torch.float32, device=base_pdf.device)
        return (1.0 / t) * base_pdf + self.log_prior(x, data)
