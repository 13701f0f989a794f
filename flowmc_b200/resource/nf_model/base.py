"""Alias of ``flowmc_b200.resource.model.nf_model.base`` (see the package docstring)."""
from ..model.nf_model import base as _m

globals().update({k: v for k, v in vars(_m).items() if not k.startswith("__")})
