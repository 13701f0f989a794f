"""NFModel -- base class of normalizing-flow resources (reference:
src/flowMC/resource/model/nf_model/base.py:14-245).

Keeps the reference's interface (``n_features``, ``data_mean``, ``data_cov``, ``forward`` /
``inverse`` / ``log_prob`` / ``sample`` / ``train`` / ``save_model`` / ``load_model``).  The
training loop (``train_step`` / ``train_epoch`` / ``train``, base.py:102-210) runs entirely on
the device: jax.random.permutation-compatible batching, one fused loss+gradient pass per batch
(``flowmc_flow_loss_grad``), optional data-parallel gradient all-reduce, and the fused
clip-by-global-norm + AdamW update (``flowmc_clip_adamw``).  The host reads ONE float per epoch
(the last batch's loss), exactly where the reference synchronises (base.py:196-200).
"""
from __future__ import annotations

from abc import abstractmethod

from ...base import Resource


class NFModel(Resource):
    _n_features: int

    @property
    def n_features(self):
        return self._n_features

    @abstractmethod
    def __init__(self):
        raise NotImplementedError

    def __call__(self, x):
        return self.forward(x)

    @abstractmethod
    def log_prob(self, x):
        raise NotImplementedError

    @abstractmethod
    def sample(self, rng_key, n_samples: int):
        raise NotImplementedError

    @abstractmethod
    def forward(self, x, key=None):
        raise NotImplementedError

    @abstractmethod
    def inverse(self, x):
        raise NotImplementedError

    def to_precision(self, precision: str = "float32"):
        """The B200 path computes in float32 only (nf_model/base.py:212-242 is experimental upstream)."""
        if precision.lower() != "float32":
            raise NotImplementedError("flowmc_b200 flows are float32")
        return self
