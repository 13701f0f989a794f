"""Alias of ``flowmc_b200.resource.kernel.base`` (see the package docstring)."""
from ..kernel.base import *  # noqa: F401,F403
from ..kernel import base as _m

globals().update({k: v for k, v in vars(_m).items() if not k.startswith("__")})
