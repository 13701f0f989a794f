"""Alias of ``flowmc_b200.resource.kernel.NF_proposal`` (see the package docstring)."""
from ..kernel.NF_proposal import *  # noqa: F401,F403
from ..kernel import NF_proposal as _m

globals().update({k: v for k, v in vars(_m).items() if not k.startswith("__")})
