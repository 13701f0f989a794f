"""flowMC quickstart (docs/tutorials/dualmoon.ipynb, BASELINE.json configs[0]) on the B200 path.

Same script as the reference tutorial with `jax.random` -> `flowmc_b200.random` and the Python target replaced by
the registered device target `dual_moon` (same formula, analytic gradient):

    5-D dual-moon, 20 chains, MALA step 0.1, MaskedCouplingRQSpline 4 layers x [32, 32] x 8 bins,
    100 local + 10 global steps per loop, 20 training + 20 production loops, 5 epochs per loop.
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from flowmc_b200 import random as jr, targets as T  # noqa: E402
from flowmc_b200.resource_strategy_bundle.RQSpline_MALA import RQSpline_MALA_Bundle  # noqa: E402
from flowmc_b200.Sampler import Sampler  # noqa: E402

n_dims, n_chains = 5, 20
rng_key = jr.PRNGKey(42)
rng_key, subkey = jr.split(rng_key)
initial_position = jr.normal(subkey, (n_chains, n_dims))

rng_key, subkey = jr.split(rng_key)
bundle = RQSpline_MALA_Bundle(
    subkey, n_chains, n_dims, T.dual_moon(), n_local_steps=100, n_global_steps=10, n_training_loops=20,
    n_production_loops=20, n_epochs=5, mala_step_size=0.1, rq_spline_hidden_units=[32, 32], rq_spline_n_bins=8,
    rq_spline_n_layers=4, learning_rate=5e-3, batch_size=5000, n_max_examples=5000, verbose=False)
sampler = Sampler(n_dims, n_chains, rng_key, resource_strategy_bundles=bundle)

torch.cuda.synchronize()
t0 = time.perf_counter()
sampler.sample(initial_position, {})
torch.cuda.synchronize()
dt = time.perf_counter() - t0

res = sampler.resources
chains = res["positions_production"].data            # [20, 2200, 5] on the GPU
ga, la = res["global_accs_production"].data, res["local_accs_production"].data
print(f"sampled {tuple(chains.shape)} in {dt:.2f} s")
print(f"local acceptance  {la[torch.isfinite(la)].mean().item():.3f}")
print(f"global acceptance {ga[torch.isfinite(ga)].mean().item():.3f}")
print(f"|x| mean {chains.norm(dim=-1).mean().item():.3f}  (dual moon: mass near |x| = 2)")
print(f"loss first/last {res['loss_buffer'].data[0].item():.3f} / {res['loss_buffer'].data[-1].item():.3f}")
