"""TrainModel strategy (reference: src/flowMC/strategy/train_model.py:11-112).

Same constructor and call contract.  The training set is selected on the device: finite rows of
the positions buffer, the last ``history_window`` of them per chain, ``n_max_examples`` drawn with
jax.random.choice-compatible indices (train_model.py:66-81).  With a chain shard set
(``set_chain_shard``) each rank gathers the rows of its own chains and an NCCL all-gather of the
per-rank row blocks gives every rank the full training set, after which ``NFModel.train`` runs
data-parallel.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import random as frandom
from .._lib import check, lib
from ..resource.buffers import Buffer
from ..resource.model.nf_model.base import NFModel
from ..resource.optimizer import Optimizer
from .base import Strategy

_u32p = C.POINTER(C.c_uint32)


class TrainModel(Strategy):
    def __repr__(self):
        return "Train " + self.model_resource

    def __init__(self, model_resource: str, data_resource: str, optimizer_resource: str,
                 loss_buffer_name: str = "", n_epochs: int = 100, batch_size: int = 64,
                 n_max_examples: int = 10000, history_window: int = 100, verbose: bool = False):
        self.model_resource = model_resource
        self.data_resource = data_resource
        self.optimizer_resource = optimizer_resource
        self.loss_buffer_name = loss_buffer_name
        self.n_epochs = n_epochs
        self.batch_size = batch_size
        self.n_max_examples = n_max_examples
        self.verbose = verbose
        self.history_window = history_window
        self.chain_shard = None   # (offset, n_chains_global, all_reduce) for multi-GPU runs
        self.shard = None
        self.last_training_data = None

    def set_chain_shard(self, offset: int, n_chains_global: int, all_reduce, shard=None):
        """``shard`` (a ``flowmc_b200.parallel.ChainShard``) enables the all-gather assembly of the training set;
        without it the rows are assembled by a sum-all-reduce of a zero-filled buffer (same result)."""
        self.chain_shard = (int(offset), int(n_chains_global), all_reduce)
        self.shard = shard

    def _assemble_all_gather(self, buf, rowmap, n_total, d, window, lo, idx, stream):
        """All-gather of per-rank row blocks.  Every rank knows the whole ``choice`` draw ``idx``; rows are
        grouped by owning rank (stable), each rank gathers the rows of its own chains into its block, one
        ``all_gather_into_tensor`` moves every block to every rank, and the blocks are put back in draw order."""
        def gather_own(my_idx, block):
            check(lib.flowmc_gather_training_rows(buf.data_ptr(), rowmap.data_ptr(), n_total, d, window, lo,
                                                  self.shard.offset, self.shard.offset + buf.shape[0],
                                                  my_idx.data_ptr(), int(my_idx.numel()), block.data_ptr(), stream))
        return self.shard.assemble_rows(idx, window, d, gather_own)

    def select_training_data(self, rng_key, buf: torch.Tensor):
        """train_model.py:66-81 -> (rng_key after the first split, training_data [n_max_examples, d])."""
        n_chains, n_total, d = buf.shape
        dev = buf.device
        stream = torch.cuda.current_stream().cuda_stream
        rowmap = torch.empty((n_chains, n_total), dtype=torch.int32, device=dev)
        counts = torch.empty(n_chains, dtype=torch.int32, device=dev)
        minmax = torch.empty(2, dtype=torch.int32, device=dev)
        offset, n_glob, all_reduce = self.chain_shard if self.chain_shard is not None else (0, n_chains, None)
        with torch.cuda.device(dev):
            check(lib.flowmc_buffer_finite_rows(buf.data_ptr(), n_chains, n_total, d, rowmap.data_ptr(),
                                                counts.data_ptr(), minmax.data_ptr(), stream))
            lo, hi = (int(v) for v in minmax.tolist())   # the reference's boolean-mask indexing syncs here too
            if all_reduce is not None:
                mm = torch.tensor([-lo, hi], dtype=torch.float32, device=dev)
                all_reduce(mm, "max")
                lo, hi = -int(mm[0].item()), int(mm[1].item())
            if lo != hi:
                raise ValueError(f"chains have different numbers of finite rows ({lo}..{hi}): the reference's "
                                 "reshape(n_chains, -1, n_dims) (train_model.py:68-70) requires them equal")
            if lo == 0:
                raise ValueError("the positions buffer holds no finite rows to train on")
            window = min(int(self.history_window), lo)
            rng_key, subkey = frandom.split(np.asarray(rng_key, dtype=np.uint32))
            subkey = np.ascontiguousarray(subkey)
            m = int(self.n_max_examples)
            idx = torch.empty(m, dtype=torch.int32, device=dev)
            check(lib.flowmc_random_choice(subkey.ctypes.data_as(_u32p), n_glob * window, m, idx.data_ptr(), stream))
            if all_reduce is not None and self.shard is not None and self.shard.world_size > 1:
                return rng_key, self._assemble_all_gather(buf, rowmap, n_total, d, window, lo, idx, stream)
            out = torch.zeros((m, d), dtype=torch.float32, device=dev) if all_reduce is not None else \
                torch.empty((m, d), dtype=torch.float32, device=dev)
            check(lib.flowmc_gather_training_rows(buf.data_ptr(), rowmap.data_ptr(), n_total, d, window, lo, offset,
                                                  offset + n_chains, idx.data_ptr(), m, out.data_ptr(), stream))
            if all_reduce is not None:
                all_reduce(out)        # disjoint rows, zeros elsewhere: the sum IS the all-gather
        return rng_key, out

    def __call__(self, rng_key, resources, initial_position, data):
        model = resources[self.model_resource]
        assert isinstance(model, NFModel), "Target resource must be a NFModel"
        data_resource = resources[self.data_resource]
        assert isinstance(data_resource, Buffer), "Data resource must be a buffer"
        optimizer = resources[self.optimizer_resource]
        assert isinstance(optimizer, Optimizer), "Optimizer resource must be an optimizer"
        rng_key, training_data = self.select_training_data(rng_key, data_resource.data)
        self.last_training_data = training_data
        rng_key, subkey = frandom.split(rng_key)

        if self.verbose:
            print("Training model")
            print(f"Training data shape: {tuple(training_data.shape)}")
            print(f"n_epochs: {self.n_epochs}")
            print(f"batch_size: {self.batch_size}")

        (rng_key, model, optim_state, loss_values) = model.train(
            rng=subkey, data=training_data, optim=optimizer.optim, state=optimizer.optim_state,
            num_epochs=self.n_epochs, batch_size=self.batch_size, verbose=self.verbose)

        if self.loss_buffer_name != "":
            loss_buffer = resources[self.loss_buffer_name]
            assert isinstance(loss_buffer, Buffer), "Loss buffer resource must be a buffer"
            loss_buffer.update_buffer(loss_values, start=loss_buffer.cursor)
            loss_buffer.cursor += len(loss_values)
            resources[self.loss_buffer_name] = loss_buffer

        optimizer.optim_state = optim_state
        resources[self.model_resource] = model
        resources[self.optimizer_resource] = optimizer
        return rng_key, resources, initial_position
