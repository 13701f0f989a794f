"""Alias of ``flowmc_b200.resource.kernel.HMC`` (see the package docstring)."""
from ..kernel.HMC import *  # noqa: F401,F403
from ..kernel import HMC as _m

globals().update({k: v for k, v in vars(_m).items() if not k.startswith("__")})
