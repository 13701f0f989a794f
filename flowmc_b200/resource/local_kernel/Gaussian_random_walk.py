"""Alias of ``flowmc_b200.resource.kernel.Gaussian_random_walk`` (see the package docstring)."""
from ..kernel.Gaussian_random_walk import *  # noqa: F401,F403
from ..kernel import Gaussian_random_walk as _m

globals().update({k: v for k, v in vars(_m).items() if not k.startswith("__")})
