"""RQSpline_MALA_Bundle (reference: src/flowMC/resource_strategy_bundle/RQSpline_MALA.py:22-274).

Same keyword arguments, resource names, strategy names and strategy order as the reference, so a
``Sampler(resource_strategy_bundles=RQSpline_MALA_Bundle(...))`` script runs unchanged apart from
the target being a ``flowmc_b200.targets`` device target.  Extra keyword ``chain_shard`` =
``flowmc_b200.parallel.ChainShard``: this process then owns only its slab of the chains (buffers
are allocated per rank) and flow training runs data-parallel.
"""
from __future__ import annotations

from .. import random as frandom
from ..resource.buffers import Buffer
from ..resource.kernel.MALA import MALA
from ..resource.kernel.NF_proposal import NFProposal
from ..resource.logPDF import LogPDF
from ..resource.model.nf_model.rqSpline import MaskedCouplingRQSpline
from ..resource.optimizer import Optimizer
from ..resource.states import State
from ..strategy.lambda_function import Lambda
from ..strategy.take_steps import TakeGroupSteps, TakeSerialSteps
from ..strategy.train_model import TrainModel
from ..strategy.update_state import UpdateState
from .base import ResourceStrategyBundle


class RQSpline_MALA_Bundle(ResourceStrategyBundle):
    """Rational-quadratic-spline flow as the global proposal + MALA as the local sampler."""

    def __repr__(self):
        return "RQSpline_MALA Bundle"

    def __init__(self, rng_key, n_chains: int, n_dims: int, logpdf, n_local_steps: int, n_global_steps: int,
                 n_training_loops: int, n_production_loops: int, n_epochs: int, mala_step_size: float = 1e-1,
                 chain_batch_size: int = 0, rq_spline_hidden_units: list = [32, 32], rq_spline_n_bins: int = 8,
                 rq_spline_n_layers: int = 4, learning_rate: float = 1e-3, batch_size: int = 10000,
                 n_max_examples: int = 10000, local_thinning: int = 1, global_thinning: int = 1,
                 n_NFproposal_batch_size: int = 10000, verbose: bool = False, chain_shard=None):
        per_loop = n_local_steps // local_thinning + n_global_steps // global_thinning
        n_training_steps = per_loop * n_training_loops
        n_production_steps = per_loop * n_production_loops
        # with a shard, this rank stores only its own chains (n_chains stays the GLOBAL count)
        n_local = n_chains if chain_shard is None else chain_shard.n_local

        def chain_buffers(tag, n_steps):
            return {
                f"positions_{tag}": Buffer(f"positions_{tag}", (n_local, n_steps, n_dims), 1),
                f"log_prob_{tag}": Buffer(f"log_prob_{tag}", (n_local, n_steps), 1),
                f"local_accs_{tag}": Buffer(f"local_accs_{tag}", (n_local, n_steps), 1),
                f"global_accs_{tag}": Buffer(f"global_accs_{tag}", (n_local, n_steps), 1),
            }

        rng_key, subkey = frandom.split(rng_key)
        model = MaskedCouplingRQSpline(n_dims, rq_spline_n_layers, rq_spline_hidden_units, rq_spline_n_bins, subkey)
        self.resources = {
            "logpdf": logpdf if isinstance(logpdf, LogPDF) else LogPDF(logpdf, n_dims=n_dims),
            **chain_buffers("training", n_training_steps),
            "loss_buffer": Buffer("loss_buffer", (n_training_loops * n_epochs,), 0),
            **chain_buffers("production", n_production_steps),
            "local_sampler": MALA(step_size=mala_step_size),
            "global_sampler": NFProposal(model, n_NFproposal_batch_size=n_NFproposal_batch_size),
            "model": model,
            "optimizer": Optimizer(model=model, learning_rate=learning_rate),
            "sampler_state": State(
                {
                    "target_positions": "positions_training",
                    "target_log_prob": "log_prob_training",
                    "target_local_accs": "local_accs_training",
                    "target_global_accs": "global_accs_training",
                    "training": True,
                },
                name="sampler_state",
            ),
        }

        local_stepper = TakeSerialSteps(
            "logpdf", "local_sampler", "sampler_state",
            ["target_positions", "target_log_prob", "target_local_accs"],
            n_local_steps, thinning=local_thinning, chain_batch_size=chain_batch_size, verbose=verbose)
        global_stepper = TakeGroupSteps(
            "logpdf", "global_sampler", "sampler_state",
            ["target_positions", "target_log_prob", "target_global_accs"],
            n_global_steps, thinning=global_thinning, chain_batch_size=chain_batch_size, verbose=verbose)
        model_trainer = TrainModel(
            "model", "positions_training", "optimizer", loss_buffer_name="loss_buffer", n_epochs=n_epochs,
            batch_size=batch_size, n_max_examples=n_max_examples, verbose=verbose)
        if chain_shard is not None:
            chain_shard.attach(local_stepper, global_stepper, model_trainer, model)

        def reset_steppers(rng_key, resources, initial_position, data):
            local_stepper.set_current_position(0)
            global_stepper.set_current_position(0)

        def update_model(rng_key, resources, initial_position, data):
            # the trained model replaces the proposal's model (RQSpline_MALA.py:215-231)
            resources["global_sampler"].model = resources["model"]

        self.strategies = {
            "local_stepper": local_stepper,
            "global_stepper": global_stepper,
            "model_trainer": model_trainer,
            "update_state": UpdateState(
                "sampler_state",
                ["target_positions", "target_log_prob", "target_local_accs", "target_global_accs", "training"],
                ["positions_production", "log_prob_production", "local_accs_production", "global_accs_production",
                 False]),
            # both steppers share one cursor so local and global samples interleave in the same buffer
            "update_global_step": Lambda(
                lambda rng_key, resources, initial_position, data:
                global_stepper.set_current_position(local_stepper.current_position)),
            "update_local_step": Lambda(
                lambda rng_key, resources, initial_position, data:
                local_stepper.set_current_position(global_stepper.current_position)),
            "reset_steppers": Lambda(reset_steppers),
            "update_model": Lambda(update_model),
        }

        training_phase = ["local_stepper", "update_global_step", "model_trainer", "update_model", "global_stepper",
                          "update_local_step"]
        production_phase = ["local_stepper", "update_global_step", "global_stepper", "update_local_step"]
        self.strategy_order = (training_phase * n_training_loops + ["reset_steppers", "update_state"]
                               + production_phase * n_production_loops)
