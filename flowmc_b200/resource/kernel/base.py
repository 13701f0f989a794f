"""ProposalBase -- the kernel plugin boundary (reference: src/flowMC/resource/kernel/base.py:9-27).

A proposal is a Resource with ``kernel(rng_key, position, log_prob, logpdf, data) ->
(position, log_prob, do_accept)``.  The reference's method evolves ONE chain and is vmapped by
the strategy; here ``kernel`` accepts either one chain (key uint32[2], position [d]) or a batch
(keys uint32[n,2], positions [n,d], log_prob [n]) and runs them through the same CUDA kernel
that ``TakeSerialSteps`` uses, with the given keys used exactly as the reference uses ``rng_key``.
"""
from __future__ import annotations

import ctypes as C
from abc import abstractmethod

import numpy as np
import torch

from ..._lib import LocalParams, check, lib
from ..base import Resource
from ..logPDF import LogPDF

_u32p = C.POINTER(C.c_uint32)


class ProposalBase(Resource):
    @abstractmethod
    def __init__(self):
        """Initialize the sampler class."""

    @abstractmethod
    def kernel(self, rng_key, position, log_prob, logpdf, data):
        """Kernel for one step in the proposal cycle."""


class LocalKernel(ProposalBase):
    """Shared host glue of the three local kernels (MALA / HMC / Gaussian random walk)."""

    KIND: int = -1
    layout_hint: int = 0
    # launch-plan overrides of flowmc_local_steps (FlowmcLocalParams.force_n_seg / slots_override): results do not
    # depend on them; tests use them to drive the time-sliced multi-launch path with small shapes
    force_n_seg: int = 0
    slots_override: int = 0

    def __init__(self):
        pass

    def _local_params(self, n_dims: int, device: torch.device) -> tuple[LocalParams, list]:
        """(params struct, tensors to keep alive during the call)."""
        raise NotImplementedError

    def kernel(self, rng_key, position, log_prob, logpdf, data):
        if not isinstance(logpdf, LogPDF):
            logpdf = LogPDF(logpdf, n_dims=int(torch.as_tensor(position).shape[-1]))
        position = torch.as_tensor(position, dtype=torch.float32)
        if not position.is_cuda:
            position = position.cuda()
        single = position.dim() == 1
        x = position.reshape(1, -1) if single else position
        x = x.contiguous()
        n, d = x.shape
        dev = x.device
        keys = np.ascontiguousarray(np.asarray(rng_key, dtype=np.uint32).reshape(-1, 2))
        if keys.shape[0] != n:
            raise ValueError(f"need one key per chain: got {keys.shape[0]} keys for {n} chains")
        keys_d = torch.from_numpy(keys.view(np.int32)).to(dev)
        lp_in = torch.as_tensor(log_prob, dtype=torch.float32, device=dev).reshape(n).contiguous()
        pos = torch.empty((n, 1, d), dtype=torch.float32, device=dev)
        lp = torch.empty((n, 1), dtype=torch.float32, device=dev)
        acc = torch.empty((n, 1), dtype=torch.float32, device=dev)
        last = torch.empty((n, d), dtype=torch.float32, device=dev)
        params, keep = self._local_params(d, dev)
        params.step_keys = keys_d.data_ptr()
        params.lp0 = lp_in.data_ptr()   # used by HMC / GRW; MALA re-evaluates logpdf(position) like MALA.py:59,75
        params.force_n_seg = 0
        pk = logpdf.target.packed_on(data, d, dev)
        dummy = np.zeros(2, np.uint32)
        out_key = np.zeros(2, np.uint32)
        with torch.cuda.device(dev):
            check(lib.flowmc_local_steps(self.KIND, logpdf.target.target_id, pk.data_ptr(),
                                         dummy.ctypes.data_as(_u32p), x.data_ptr(), pos.data_ptr(), lp.data_ptr(),
                                         acc.data_ptr(), 1, 0, n, d, 1, 1, 0, n, C.byref(params),
                                         out_key.ctypes.data_as(_u32p), last.data_ptr(),
                                         torch.cuda.current_stream().cuda_stream))
        do_accept = acc[:, 0] > 0.5
        if single:
            return last[0], lp[0, 0], do_accept[0]
        return last, lp[:, 0], do_accept
