"""TakeSteps / TakeSerialSteps / TakeGroupSteps (reference: src/flowMC/strategy/take_steps.py:14-206).

Same constructor, attributes and call contract.  Where the reference does
``filter_jit(filter_vmap(sample))`` + three functional buffer updates per call
(take_steps.py:127-142), ``TakeSerialSteps`` makes ONE C-ABI call (``flowmc_local_steps``): a
persistent CUDA kernel runs all ``n_steps`` for every chain and stores the thinned positions,
log-probs and accept flags directly into the buffers at ``current_position``.
``chain_batch_size`` is accepted and ignored (the reference's chain micro-batching,
take_steps.py:106-125, exists only to bound vmap memory; results do not depend on it).

Multi-GPU: chains are sharded by ``flowmc_b200.parallel.ChainShard`` -- each process owns the
global chains [offset, offset+n) and the per-chain keys are taken from the *global* split, so
any sharding reproduces the single-GPU chains bit for bit with no communication.
"""
from __future__ import annotations

import ctypes as C
from abc import abstractmethod

import numpy as np
import torch

from .._lib import check, lib
from ..resource.buffers import Buffer, clamp_start
from ..resource.kernel.base import LocalKernel, ProposalBase
from ..resource.logPDF import LogPDF
from ..resource.states import State
from .base import Strategy

_u32p = C.POINTER(C.c_uint32)


def _to_device(x) -> torch.Tensor:
    x = torch.as_tensor(x, dtype=torch.float32)
    if not x.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError("flowmc_b200 needs a CUDA device (there is no CPU fallback)")
        x = x.cuda()
    return x.contiguous()


class TakeSteps(Strategy):
    def __init__(self, logpdf_name: str, kernel_name: str, state_name: str, buffer_names: list,
                 n_steps: int, thinning: int = 1, chain_batch_size: int = 0, verbose: bool = False):
        self.logpdf_name = logpdf_name
        self.kernel_name = kernel_name
        self.state_name = state_name
        self.buffer_names = buffer_names
        self.n_steps = n_steps
        self.current_position = 0
        self.thinning = thinning
        self.chain_batch_size = chain_batch_size
        self.verbose = verbose
        # global chain shard owned by this process: (offset, n_chains_global) or None = all chains
        self.chain_shard = None
        self._workspace = None

    def set_current_position(self, current_position: int):
        self.current_position = current_position

    def set_chain_shard(self, offset: int, n_chains_global: int):
        self.chain_shard = (int(offset), int(n_chains_global))

    @abstractmethod
    def sample(self, kernel, rng_key, initial_position, logpdf, data):
        raise NotImplementedError

    def _resolve(self, resources):
        assert isinstance(state_resource := resources[self.state_name], State), "State resource must be a State"
        names = []
        for i, what in enumerate(("Position", "Log probability", "Acceptance")):
            assert isinstance(nm := state_resource.data[self.buffer_names[i]], str), \
                f"{what} buffer resource name must be a string"
            names.append(nm)
        bufs = []
        for nm, what in zip(names, ("Position", "Log probability", "Acceptance")):
            assert isinstance(b := resources[nm], Buffer), f"{what} buffer resource must be a Buffer"
            bufs.append(b)
        return bufs

    def _n_out(self) -> int:
        return len(range(0, self.n_steps, self.thinning))

    def __call__(self, rng_key, resources, initial_position, data):
        position_buffer, log_prob_buffer, acceptance_buffer = self._resolve(resources)
        kernel = resources[self.kernel_name]
        logpdf = resources[self.logpdf_name]
        x0 = _to_device(initial_position)
        if x0.dim() == 1:
            x0 = x0.reshape(1, -1)
        n_out = self._n_out()
        # dynamic_update_slice clamps the start so the update fits (buffers.py:39-41)
        start = clamp_start(self.current_position, n_out, position_buffer.data.shape[1])
        rng_key, last = self.sample(kernel, rng_key, x0, logpdf, data,
                                    (position_buffer, log_prob_buffer, acceptance_buffer), start)
        self.current_position += self.n_steps // self.thinning
        return rng_key, resources, last


class TakeSerialSteps(TakeSteps):
    """Takes ``n_steps`` dependent steps of a local kernel for every chain (one kernel launch)."""

    def sample(self, kernel, rng_key, x0, logpdf, data, buffers, start):
        if not isinstance(kernel, LocalKernel):
            raise TypeError("TakeSerialSteps drives MALA / HMC / GaussianRandomWalk kernels")
        assert isinstance(logpdf, LogPDF), "logpdf resource must be a LogPDF"
        pos_b, lp_b, acc_b = buffers
        n, d = x0.shape
        dev = x0.device
        for b in (pos_b, lp_b, acc_b):
            if b.data.device != dev or not b.data.is_contiguous() or b.data.shape[0] != n:
                raise ValueError(f"buffer {b.name} must be a contiguous tensor on {dev} with {n} chains")
        if pos_b.data.shape[2] != d or lp_b.data.shape[1] != pos_b.data.shape[1] \
                or acc_b.data.shape[1] != pos_b.data.shape[1]:
            raise ValueError("position / log-prob / acceptance buffers have inconsistent shapes")
        n_total = pos_b.data.shape[1]
        offset, n_glob = self.chain_shard if self.chain_shard is not None else (0, n)
        params, keep = kernel._local_params(d, dev)
        ws_bytes = int(lib.flowmc_local_steps_workspace_bytes(n, d, params.layout_hint))
        if self._workspace is None or self._workspace.numel() < ws_bytes or self._workspace.device != dev:
            self._workspace = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=dev)
        params.workspace = self._workspace.data_ptr()
        params.workspace_bytes = self._workspace.numel()
        pk = logpdf.target.packed_on(data, d, dev)
        key = np.ascontiguousarray(rng_key, dtype=np.uint32)
        key_out = np.zeros(2, np.uint32)
        last = torch.empty((n, d), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.flowmc_local_steps(kernel.KIND, logpdf.target.target_id, pk.data_ptr(),
                                         key.ctypes.data_as(_u32p), x0.data_ptr(), pos_b.data.data_ptr(),
                                         lp_b.data.data_ptr(), acc_b.data.data_ptr(), n_total, start, n, d,
                                         self.n_steps, self.thinning, offset, n_glob, C.byref(params),
                                         key_out.ctypes.data_as(_u32p), last.data_ptr(),
                                         torch.cuda.current_stream().cuda_stream))
        return key_out, last


class TakeGroupSteps(TakeSteps):
    """Takes ``n_steps`` independent proposals at once (normalizing-flow global steps)."""

    def sample(self, kernel, rng_key, x0, logpdf, data, buffers, start):
        if not hasattr(kernel, "group_steps"):
            raise TypeError("TakeGroupSteps drives kernels with a fused group step (NFProposal)")
        offset, n_glob = self.chain_shard if self.chain_shard is not None else (0, x0.shape[0])
        return kernel.group_steps(rng_key, x0, logpdf, data, buffers, start, self.n_steps, self.thinning,
                                  offset, n_glob)
