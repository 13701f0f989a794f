"""RQSpline_MALA_PT_Bundle (reference: src/flowMC/resource_strategy_bundle/RQSpline_MALA_PT.py:24-344).

The RQSpline_MALA bundle plus a parallel-tempering step in front of every local step: same keyword arguments,
resource names ("tempered_logpdf", "tempered_positions", "temperatures"), strategy names and strategy order as the
reference.  ``logprior`` is a ``flowmc_b200.resource.logPDF.BoxQuadraticPrior`` (or None for the reference's default,
the flat prior 0): the prior runs inside the tempered sampling kernel.
"""
from __future__ import annotations

import torch

from ..resource.buffers import Buffer
from ..resource.logPDF import TemperedPDF
from ..strategy.lambda_function import Lambda
from ..strategy.parallel_tempering import ParallelTempering
from .RQSpline_MALA import RQSpline_MALA_Bundle


class RQSpline_MALA_PT_Bundle(RQSpline_MALA_Bundle):
    """Rational-quadratic-spline flow as the global proposal, MALA as the local sampler, and parallel tempering."""

    def __repr__(self):
        return "RQSpline MALA PT Bundle"

    def __init__(self, rng_key, n_chains: int, n_dims: int, logpdf, n_local_steps: int, n_global_steps: int,
                 n_training_loops: int, n_production_loops: int, n_epochs: int, mala_step_size: float = 1e-1,
                 chain_batch_size: int = 0, rq_spline_hidden_units: list = [32, 32], rq_spline_n_bins: int = 8,
                 rq_spline_n_layers: int = 4, learning_rate: float = 1e-3, batch_size: int = 10000,
                 n_max_examples: int = 10000, local_thinning: int = 1, global_thinning: int = 1,
                 n_NFproposal_batch_size: int = 10000, history_window: int = 100, n_temperatures: int = 5,
                 max_temperature: float = 5.0, n_tempered_steps: int = -1, logprior=None, verbose: bool = False,
                 chain_shard=None):
        super().__init__(rng_key, n_chains, n_dims, logpdf, n_local_steps, n_global_steps, n_training_loops,
                         n_production_loops, n_epochs, mala_step_size=mala_step_size,
                         chain_batch_size=chain_batch_size, rq_spline_hidden_units=rq_spline_hidden_units,
                         rq_spline_n_bins=rq_spline_n_bins, rq_spline_n_layers=rq_spline_n_layers,
                         learning_rate=learning_rate, batch_size=batch_size, n_max_examples=n_max_examples,
                         local_thinning=local_thinning, global_thinning=global_thinning,
                         n_NFproposal_batch_size=n_NFproposal_batch_size, verbose=verbose, chain_shard=chain_shard)
        self.strategies["model_trainer"].history_window = history_window
        n_local = n_chains if chain_shard is None else chain_shard.n_local

        # the resources of the parallel tempering (RQSpline_MALA_PT.py:113-129)
        tempered_logpdf = TemperedPDF(self.resources["logpdf"], logprior, n_dims=n_dims, n_temps=n_temperatures)
        tempered_positions = Buffer("tempered_positions", (n_local, n_temperatures - 1, n_dims), 2)
        temperatures = Buffer("temperature", (n_temperatures,), 0)
        temperatures.update_buffer(torch.linspace(1.0, max_temperature, n_temperatures))
        self.resources.update({"tempered_logpdf": tempered_logpdf, "tempered_positions": tempered_positions,
                               "temperatures": temperatures})

        if n_tempered_steps <= 0:
            print("n_tempered_steps value is not valid. Setting to n_local_steps")
            n_tempered_steps = n_local_steps
        parallel_tempering_strat = ParallelTempering(
            n_steps=n_tempered_steps, tempered_logpdf_name="tempered_logpdf", kernel_name="local_sampler",
            tempered_buffer_names=["tempered_positions", "temperatures"], state_name="sampler_state", verbose=verbose)
        if chain_shard is not None:
            parallel_tempering_strat.set_chain_shard(chain_shard.offset, chain_shard.n_chains_global,
                                                     chain_shard.all_reduce)

        def initialize_tempered_positions(rng_key, resources, initial_position, data):
            # every rung starts at the chain's initial position (RQSpline_MALA_PT.py:283-287)
            x0 = torch.as_tensor(initial_position, dtype=torch.float32, device=tempered_positions.data.device)
            tempered_positions.update_buffer(x0[:, None, :].repeat(1, n_temperatures - 1, 1))

        self.strategies.update({"parallel_tempering": parallel_tempering_strat,
                                "initialize_tempered_positions": Lambda(initialize_tempered_positions)})

        training_phase = ["parallel_tempering", "local_stepper", "update_global_step", "model_trainer", "update_model",
                          "global_stepper", "update_local_step"]
        production_phase = ["parallel_tempering", "local_stepper", "update_global_step", "global_stepper",
                            "update_local_step"]
        self.strategy_order = (["initialize_tempered_positions"] + training_phase * n_training_loops
                               + ["reset_steppers", "update_state"] + production_phase * n_production_loops)
