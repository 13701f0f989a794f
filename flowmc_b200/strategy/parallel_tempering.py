"""ParallelTempering (reference: src/flowMC/strategy/parallel_tempering.py:13-436).

Same constructor and call contract.  One call = (1) ``n_steps`` of the local kernel on every (chain, temperature)
pair of the tempered density, (2) one sweep of neighbour exchanges up the temperature ladder, (3) while
``state.data["training"]``: store the tempered positions and adapt the temperatures from the exchange acceptance.

On the B200 path (1) is ONE launch of the persistent local-step kernel over ``n_chains * n_temps`` virtual chains
(``flowmc_local_steps(FLOWMC_KERNEL_{MALA,HMC,GRW}_TEMPERED)``: per-chain inverse temperature, prior evaluated in the
kernel, explicit per-(chain, temperature) keys ``split(split(subkey, n_chains)[c], n_temps)[t]``), (2) is
``flowmc_pt_exchange`` (one thread per chain walks the ladder), and (3) is the reference's arithmetic on ``n_temps``
numbers on the host.  Any of the three local kernels can be the tempered kernel, as in the reference, which hands
whatever ``ProposalBase`` the resources name to ``_individual_step`` (parallel_tempering.py:74,135-289; its PT bundle
uses MALA, RQSpline_MALA_PT.py).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import random as frandom
from .._lib import LocalParams, check, lib
from ..resource.buffers import Buffer
from ..resource.kernel.base import LocalKernel, ProposalBase
from ..resource.logPDF import TemperedPDF
from ..resource.states import State
from .base import Strategy

_u32p = C.POINTER(C.c_uint32)
_TEMPERED_KIND_OFFSET = 3      # FLOWMC_KERNEL_{MALA,HMC,GRW}_TEMPERED = 3 + the plain kind


class ParallelTempering(Strategy):
    """Sample a tempered PDF with one exchange step (see the reference docstring, parallel_tempering.py:14-23)."""

    def __init__(self, n_steps: int, tempered_logpdf_name: str, kernel_name: str, tempered_buffer_names: list,
                 state_name: str, verbose: bool = False):
        self.n_steps = n_steps
        self.tempered_logpdf_name = tempered_logpdf_name
        self.kernel_name = kernel_name
        self.tempered_buffer_names = tempered_buffer_names
        self.verbose = verbose
        self.state_name = state_name
        # global chain shard owned by this process: (offset, n_chains_global) or None = all chains
        self.chain_shard = None
        self.all_reduce = None

    def set_chain_shard(self, offset: int, n_chains_global: int, all_reduce=None):
        """``all_reduce(tensor)`` (sum over ranks, in place): makes the temperature adaptation use the exchange
        acceptance of ALL chains, so every rank adapts the same ladder as a single-GPU run."""
        self.chain_shard = (int(offset), int(n_chains_global))
        self.all_reduce = all_reduce

    def __call__(self, rng_key, resources, initial_position, data):
        rng_key, subkey = frandom.split(rng_key)                       # parallel_tempering.py:73 (subkey unused there too)
        assert isinstance(kernel := resources[self.kernel_name], ProposalBase)
        assert isinstance(tempered_logpdf := resources[self.tempered_logpdf_name], TemperedPDF)
        assert isinstance(tempered_positions := resources[self.tempered_buffer_names[0]], Buffer)
        assert isinstance(temperatures := resources[self.tempered_buffer_names[1]], Buffer)
        assert isinstance(state := resources[self.state_name], State)

        x0 = torch.as_tensor(initial_position, dtype=torch.float32, device=tempered_positions.data.device)
        positions = torch.cat([x0[:, None, :], tempered_positions.data], dim=1).contiguous()   # [n_chains, n_temps, d]

        # take individual steps (:91-101)
        rng_key, subkey = frandom.split(rng_key)
        positions, log_probs, do_accepts = self._ensemble_steps(kernel, subkey, positions, tempered_logpdf,
                                                                temperatures.data, data)
        if self.verbose:
            print("Mean acceptance of individual steps in PT: " + str(float(do_accepts.mean())))

        # exchange between temperatures (:110-115)
        rng_key, subkey = frandom.split(rng_key)
        positions, log_probs, do_accepts = self._exchange(subkey, positions, tempered_logpdf, temperatures.data, data)
        if self.verbose:
            print("Mean acceptance of exchange steps in PT: " + str(float(do_accepts.mean())))

        if state.data["training"]:                                      # :123-132
            tempered_positions.update_buffer(positions[:, 1:], 0)
            temperatures.update_buffer(self._adapt_temperature(temperatures.data, do_accepts), 0)
        return rng_key, resources, positions[:, 0].contiguous()

    # ---- (1) individual steps: _ensemble_step / _individal_step / _individual_step_body (:134-290) ------------
    def _ensemble_steps(self, kernel, subkey, positions, logpdf: TemperedPDF, temperatures, data):
        """``subkey``: the key whose ``split(subkey, n_chains)`` the reference vmaps over.  Returns the final positions
        [n_chains, n_temps, d], final tempered log-probs [n_chains, n_temps], accept flags [n_chains, n_temps, n_steps]."""
        if not isinstance(kernel, LocalKernel) or kernel.KIND not in (0, 1, 2):
            raise NotImplementedError("flowmc_b200 ParallelTempering runs MALA, HMC or GaussianRandomWalk as the tempered "
                                      "kernel")
        n, n_temps, d = positions.shape
        dev = positions.device
        offset, n_glob = self.chain_shard if self.chain_shard is not None else (0, n)
        chain_keys = frandom.split(subkey, n_glob)[offset:offset + n]                       # :98
        keys = frandom.split_each(chain_keys, n_temps).reshape(n * n_temps, 2)               # :283
        keys_d = torch.from_numpy(np.ascontiguousarray(keys).view(np.int32)).to(dev)
        temps = torch.as_tensor(temperatures, dtype=torch.float32, device=dev).reshape(n_temps)
        beta = (1.0 / temps).repeat(n).contiguous()                                         # logPDF.py:106
        nv, T_ = n * n_temps, int(self.n_steps)
        x0 = positions.reshape(nv, d).contiguous()
        if T_ == 0:
            return positions, logpdf.tempered_log_pdf(temps, positions, data), torch.zeros((n, n_temps, 0), device=dev)
        pos = torch.empty((nv, T_, d), dtype=torch.float32, device=dev)
        lp = torch.empty((nv, T_), dtype=torch.float32, device=dev)
        acc = torch.empty((nv, T_), dtype=torch.float32, device=dev)
        last = torch.empty((nv, d), dtype=torch.float32, device=dev)
        p, keep = kernel._local_params(d, dev)           # step size, HMC mass constants, layout hint
        p.chain_keys = keys_d.data_ptr()
        p.beta = beta.data_ptr()
        prior_d = None
        if not logpdf.log_prior.is_flat():
            prior_d = torch.from_numpy(logpdf.log_prior.packed(d)).to(dev)
            p.prior = prior_d.data_ptr()
        pk = logpdf.target.packed_on(data, d, dev)
        dummy_key = np.zeros(2, np.uint32)
        key_out = np.zeros(2, np.uint32)
        with torch.cuda.device(dev):
            check(lib.flowmc_local_steps(_TEMPERED_KIND_OFFSET + kernel.KIND, logpdf.target.target_id, pk.data_ptr(),
                                         dummy_key.ctypes.data_as(_u32p), x0.data_ptr(), pos.data_ptr(), lp.data_ptr(),
                                         acc.data_ptr(), T_, 0, nv, d, T_, 1, 0, nv, C.byref(p),
                                         key_out.ctypes.data_as(_u32p), last.data_ptr(),
                                         torch.cuda.current_stream().cuda_stream))
        return last.reshape(n, n_temps, d), lp[:, -1].reshape(n, n_temps), acc.reshape(n, n_temps, T_)

    # ---- (2) exchange (:291-398) ---------------------------------------------------------------------------------
    def _exchange(self, subkey, positions, logpdf: TemperedPDF, temperatures, data):
        """Returns positions [n_chains, n_temps, d], UNtempered log-probs [n_chains, n_temps] (both after the swaps) and
        accept flags [n_chains, n_temps - 1]."""
        n, n_temps, d = positions.shape
        dev = positions.device
        positions = positions.contiguous().clone()
        log_probs = logpdf(positions.reshape(n * n_temps, d), data).reshape(n, n_temps).contiguous()   # :381
        temps = torch.as_tensor(temperatures, dtype=torch.float32, device=dev).reshape(n_temps).contiguous()
        acc = torch.zeros((n, max(n_temps - 1, 0)), dtype=torch.float32, device=dev)
        offset, n_glob = self.chain_shard if self.chain_shard is not None else (0, n)
        key = np.ascontiguousarray(subkey, dtype=np.uint32)
        with torch.cuda.device(dev):
            check(lib.flowmc_pt_exchange(key.ctypes.data_as(_u32p), offset, n_glob, n, n_temps, d, positions.data_ptr(),
                                         log_probs.data_ptr(), temps.data_ptr(), acc.data_ptr(),
                                         torch.cuda.current_stream().cuda_stream))
        return positions, log_probs, acc

    # ---- (3) temperature adaptation (:400-436) ----------------------------------------------------------------------
    def _adapt_temperature(self, temperatures, do_accept):
        """float32 arithmetic of the reference on the n_temps ladder values (host)."""
        t = np.asarray(torch.as_tensor(temperatures).detach().cpu(), dtype=np.float32)
        acc = torch.as_tensor(do_accept, dtype=torch.float32)
        n_chains = acc.shape[0]
        if self.chain_shard is not None and self.all_reduce is not None:
            # mean over ALL chains (parallel_tempering.py:421): accept flags are 0 / 1, so the per-rung counts add
            # exactly in float32 (< 2^24 chains) and every rank gets the single-GPU acceptance rate
            counts = acc.sum(dim=0)
            self.all_reduce(counts)
            n_chains = self.chain_shard[1]
            acceptance_rate = (counts / np.float32(n_chains)).detach().cpu().numpy().astype(np.float32)
        else:
            acceptance_rate = acc.mean(dim=0).detach().cpu().numpy().astype(np.float32)
        damping_factor = (np.float32(100.0 / n_chains) * (acceptance_rate[:-1] - acceptance_rate[1:])).astype(np.float32)
        new_t = t.copy()
        for i in range(1, t.shape[0] - 1):
            new_t[i] = new_t[i - 1] + (t[i] - t[i - 1]) * np.exp(damping_factor[i - 1], dtype=np.float32)
        return torch.from_numpy(new_t).to(torch.as_tensor(temperatures).device)
