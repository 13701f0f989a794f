"""NFProposal -- independence-MH global steps with flow proposals (reference:
src/flowMC/resource/kernel/NF_proposal.py:15-184).

Same constructor and ``kernel`` contract.  ``TakeGroupSteps`` calls ``group_steps``, which is ONE
C-ABI call (``flowmc_nf_global_steps``): flow sampling (inverse pass), flow log-probs (forward
pass), target log-probs and the sequential accept scan run on the device and the thinned samples
are stored straight into the sampler buffers.  ``n_NFproposal_batch_size`` exists in the
reference to bound memory; here it only selects the proposals' key schedule (batched vs
un-batched branch of ``sample_flow``, NF_proposal.py:135-172), which is reproduced exactly.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ..._lib import GlobalParams, check, lib
from ..logPDF import LogPDF
from ..model.nf_model.base import NFModel
from .base import ProposalBase

_u32p = C.POINTER(C.c_uint32)


class NFProposal(ProposalBase):
    model: NFModel
    n_batch_size: int

    def __repr__(self):
        return "NF proposal with " + self.model.__repr__()

    def __init__(self, model: NFModel, n_NFproposal_batch_size: int = 100):
        self.model = model
        self.n_batch_size = n_NFproposal_batch_size
        self._workspace = None

    def _run_generic(self, key, x0, logpdf, data, bufs, n_total, start, n_steps, thinning, offset, n_glob,
                     chain_keys=None, lp0=None):
        """Any NFModel with ``log_prob`` and ``sample_rows`` (RealNVP): the passes of NF_proposal.py:27-128 as
        separate launches -- target and flow log-probs of the current positions, proposals drawn with the
        reference's per-chain key schedule, their flow and target log-probs, then the sequential accept scan
        (``flowmc_nf_accept_scan``) writing the thinned samples into the buffers."""
        from ... import random as frandom
        pos, lp, acc = bufs
        n, d = x0.shape
        dev = x0.device
        key = np.ascontiguousarray(key, dtype=np.uint32)
        key_out, subkey = frandom.split(key)                                              # take_steps.py:71
        if chain_keys is None:
            ck = frandom.split(subkey, n_glob)[offset:offset + n]                         # take_steps.py:72
        else:
            ck = np.ascontiguousarray(chain_keys.cpu().numpy().view(np.uint32).reshape(n, 2))
        sub = frandom.split_each(ck, 2)[:, 1]                                             # NF_proposal.py:41
        if n_steps > self.n_batch_size:                                                   # NF_proposal.py:135-163
            n_batch = -(-n_steps // self.n_batch_size)
            n_sample = -(-n_steps // n_batch)
            keys_b = np.zeros((n, n_batch, 2), np.uint32)
            carry = sub
            for b in range(n_batch):
                two = frandom.split_each(carry, 2)
                carry, keys_b[:, b] = two[:, 0], two[:, 1]
            keys_d = torch.from_numpy(np.ascontiguousarray(keys_b.reshape(-1, 2)).view(np.int32)).to(dev)
            props = self.model.sample_rows(keys_d, n_sample).reshape(n, n_batch * n_sample, d)[:, :n_steps].contiguous()
        else:
            keys_d = torch.from_numpy(np.ascontiguousarray(sub).view(np.int32)).to(dev)
            props = self.model.sample_rows(keys_d, n_steps).reshape(n, n_steps, d)
        lp_nf_cur = self.model.log_prob(x0).contiguous()                                  # NF_proposal.py:44
        lp_nf_prop = self.model.log_prob(props.reshape(-1, d)).contiguous()               # NF_proposal.py:167
        lp_prop = logpdf(props.reshape(-1, d), data).contiguous()                         # NF_proposal.py:50-89
        lp0 = logpdf(x0, data).contiguous() if lp0 is None else lp0                       # take_steps.py:201
        ck_d = torch.from_numpy(np.ascontiguousarray(ck).view(np.int32)).to(dev)
        last = torch.empty((n, d), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.flowmc_nf_accept_scan(ck_d.data_ptr(), n, d, n_steps, thinning, x0.data_ptr(), lp0.data_ptr(),
                                            lp_nf_cur.data_ptr(), props.data_ptr(), lp_prop.data_ptr(),
                                            lp_nf_prop.data_ptr(), pos.data_ptr(), lp.data_ptr(), acc.data_ptr(),
                                            n_total, start, last.data_ptr(), torch.cuda.current_stream().cuda_stream))
        return key_out, last

    def _run(self, key, x0, logpdf, data, bufs, n_total, start, n_steps, thinning, offset, n_glob,
             chain_keys=None, lp0=None):
        assert isinstance(logpdf, LogPDF), "logpdf resource must be a LogPDF"
        pos, lp, acc = bufs
        n, d = x0.shape
        dev = x0.device
        if d != self.model.n_features:
            raise ValueError(f"flow has {self.model.n_features} features, chains have {d}")
        if not hasattr(self.model.desc, "num_bins"):      # not the spline flow: no fused proposal kernel
            return self._run_generic(key, x0, logpdf, data, bufs, n_total, start, n_steps, thinning, offset, n_glob,
                                     chain_keys, lp0)
        self.model.prepare()
        ws_bytes = int(lib.flowmc_nf_global_steps_workspace_bytes(n, d, n_steps))
        if self._workspace is None or self._workspace.numel() < ws_bytes or self._workspace.device != dev:
            self._workspace = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
        gp = GlobalParams()
        gp.n_batch_size = int(self.n_batch_size)
        gp.chain_keys = chain_keys.data_ptr() if chain_keys is not None else None
        gp.lp0 = lp0.data_ptr() if lp0 is not None else None
        gp.workspace = self._workspace.data_ptr()
        gp.workspace_bytes = self._workspace.numel()
        pk = logpdf.target.packed_on(data, d, dev)
        key = np.ascontiguousarray(key, dtype=np.uint32)
        key_out = np.zeros(2, np.uint32)
        last = torch.empty((n, d), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.flowmc_nf_global_steps(C.byref(self.model.desc), self.model.params.data_ptr(),
                                             logpdf.target.target_id, pk.data_ptr(), key.ctypes.data_as(_u32p),
                                             x0.data_ptr(), pos.data_ptr(), lp.data_ptr(), acc.data_ptr(),
                                             n_total, start, n, n_steps, thinning, offset, n_glob, C.byref(gp),
                                             key_out.ctypes.data_as(_u32p), last.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream))
        return key_out, last

    def group_steps(self, rng_key, x0, logpdf, data, buffers, start, n_steps, thinning, offset, n_glob):
        """TakeGroupSteps' fused call: writes into the Buffers at ``start``; returns (new key, positions[:, -1])."""
        pos_b, lp_b, acc_b = buffers
        n, d = x0.shape
        for b in (pos_b, lp_b, acc_b):
            if b.data.device != x0.device or not b.data.is_contiguous() or b.data.shape[0] != n:
                raise ValueError(f"buffer {b.name} must be a contiguous tensor on {x0.device} with {n} chains")
        return self._run(rng_key, x0, logpdf, data, (pos_b.data, lp_b.data, acc_b.data), pos_b.data.shape[1], start,
                         n_steps, thinning, offset, n_glob)

    def kernel(self, rng_key, position, log_prob, logpdf, data):
        """ProposalBase contract (NF_proposal.py:27-128): ``data["n_steps"]`` proposals for one chain
        (key uint32[2], position [d]) or a batch (keys [n,2], positions [n,d], log_prob [n])."""
        n_steps = int(data["n_steps"])
        tdata = {k: v for k, v in data.items() if k != "n_steps"} or None
        if not isinstance(logpdf, LogPDF):
            logpdf = LogPDF(logpdf, n_dims=int(torch.as_tensor(position).shape[-1]))
        position = torch.as_tensor(position, dtype=torch.float32)
        if not position.is_cuda:
            position = position.to(self.model.params.device)
        single = position.dim() == 1
        x = (position.reshape(1, -1) if single else position).contiguous()
        n, d = x.shape
        dev = x.device
        keys = np.ascontiguousarray(np.asarray(rng_key, dtype=np.uint32).reshape(-1, 2))
        if keys.shape[0] != n:
            raise ValueError(f"need one key per chain: got {keys.shape[0]} keys for {n} chains")
        keys_d = torch.from_numpy(keys.view(np.int32)).to(dev)
        lp_in = torch.as_tensor(log_prob, dtype=torch.float32, device=dev).reshape(n).contiguous()
        pos = torch.empty((n, n_steps, d), dtype=torch.float32, device=dev)
        lp = torch.empty((n, n_steps), dtype=torch.float32, device=dev)
        acc = torch.empty((n, n_steps), dtype=torch.float32, device=dev)
        self._run(np.zeros(2, np.uint32), x, logpdf, tdata, (pos, lp, acc), n_steps, 0, n_steps, 1, 0, n,
                  chain_keys=keys_d, lp0=lp_in)
        do_accept = acc > 0.5
        if single:
            return pos[0], lp[0], do_accept[0]
        return pos, lp, do_accept

    def sample_flow(self, rng_key, n_steps: int):
        """NF_proposal.py:130-172 for one key: (proposals [n_steps, d], flow log-probs [n_steps])."""
        from ... import random as frandom
        m = self.model
        if n_steps > self.n_batch_size:
            n_batch = -(-n_steps // self.n_batch_size)
            n_sample = -(-n_steps // n_batch)
            key = np.asarray(rng_key, dtype=np.uint32)
            xs = []
            for _ in range(n_batch):
                key, sub = frandom.split(key)
                xs.append(m.sample(sub, n_sample))
            x = torch.cat(xs)[:n_steps]
        else:
            x = m.sample(rng_key, n_steps)
        return x, m.log_prob(x)

    def print_parameters(self):
        raise NotImplementedError

    def save_resource(self, path):
        raise NotImplementedError

    def load_resource(self, path):
        raise NotImplementedError
