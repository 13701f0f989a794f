"""Alias of ``flowmc_b200.resource.kernel.MALA`` (see the package docstring)."""
from ..kernel.MALA import *  # noqa: F401,F403
from ..kernel import MALA as _m

globals().update({k: v for k, v in vars(_m).items() if not k.startswith("__")})
