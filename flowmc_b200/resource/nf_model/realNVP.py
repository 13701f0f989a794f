"""Alias of ``flowmc_b200.resource.model.nf_model.realNVP`` (see the package docstring)."""
from ..model.nf_model import realNVP as _m

globals().update({k: v for k, v in vars(_m).items() if not k.startswith("__")})
