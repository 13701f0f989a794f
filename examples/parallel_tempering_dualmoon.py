"""The parallel-tempering tutorial (docs/tutorials/parallel_tempering.ipynb) on the B200 path: the dual-moon target
sampled with the RQSpline + MALA + parallel-tempering bundle, after an Adam pre-optimisation of the start points.

Differences from the reference script: `jax.random` -> `flowmc_b200.random`, the target is the registered device
function `dual_moon`, and the prior is a `BoxQuadraticPrior` (here a wide Gaussian in a box) because it is evaluated
inside the tempered sampling kernel.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from flowmc_b200 import random as jr, targets as T  # noqa: E402
from flowmc_b200.resource.logPDF import BoxQuadraticPrior  # noqa: E402
from flowmc_b200.resource_strategy_bundle.RQSpline_MALA_PT import RQSpline_MALA_PT_Bundle  # noqa: E402
from flowmc_b200.Sampler import Sampler  # noqa: E402
from flowmc_b200.strategy.optimization import AdamOptimization  # noqa: E402

n_dims, n_chains = 5, 64
rng_key = jr.PRNGKey(42)
rng_key, subkey = jr.split(rng_key)
initial_position = jr.normal(subkey, (n_chains, n_dims)) * 3.0

# a few noisy Adam steps pull far-away chains towards the mass before sampling (strategy/optimization.py)
target = T.dual_moon()
rng_key, initial_position, logp = AdamOptimization(
    target, n_steps=50, learning_rate=5e-2, noise_level=1.0, bounds=np.array([[-10.0, 10.0]])).optimize(
        rng_key, None, initial_position, {})
print(f"after Adam: mean log-density {float(logp.mean()):.2f}")

rng_key, subkey = jr.split(rng_key)
bundle = RQSpline_MALA_PT_Bundle(
    subkey, n_chains, n_dims, target, n_local_steps=50, n_global_steps=10, n_training_loops=10, n_production_loops=10,
    n_epochs=5, mala_step_size=0.1, rq_spline_hidden_units=[32, 32], rq_spline_n_bins=8, rq_spline_n_layers=4,
    learning_rate=5e-3, batch_size=2000, n_max_examples=4000, n_temperatures=5, max_temperature=10.0,
    n_tempered_steps=5, logprior=BoxQuadraticPrior(c=1.0 / (2 * 10.0 ** 2), lower=-10.0, upper=10.0))

sampler = Sampler(n_dims, n_chains, rng_key, resource_strategy_bundles=bundle)
sampler.sample(initial_position, {})
torch.cuda.synchronize()

chains = sampler.resources["positions_production"].data          # [n_chains, 600, n_dims] on the GPU
temps = sampler.resources["temperatures"].data.cpu().numpy()
gacc = sampler.resources["global_accs_production"].data
print("temperature ladder after adaptation:", np.round(temps, 3))
print(f"production samples: {tuple(chains.shape)}, mean |x| = {float(chains.norm(dim=-1).mean()):.3f}, "
      f"global acceptance = {float(gacc[torch.isfinite(gacc)].mean()):.3f}")
