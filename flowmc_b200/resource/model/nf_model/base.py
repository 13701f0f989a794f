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

import os

import ctypes as C
from abc import abstractmethod

import numpy as np
import torch

from .... import random as frandom
from ...._lib import check, lib
from ...base import Resource
from ....tracing import nvtx_range

_u32p = C.POINTER(C.c_uint32)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class _TrainScratch:
    """Device scratch of one ``train`` call (allocated once, reused by every step)."""

    def __init__(self, model, n_rows: int, batch_rows: int):
        dev = model.params.device
        d = model.desc
        # flat gradient + the loss in ONE buffer: the data-parallel step all-reduces both with a single collective.
        # With a peer group (ranks of one box over NVLink) the buffer IS this rank's exchange block and the collective
        # happens inside the optimiser kernel (flowmc_dp_reduce_adamw); otherwise NCCL all-reduce + flowmc_clip_adamw.
        self.peer = None
        if model.dp is not None and len(model.dp) > 4 and model.dp[1] > 1:
            self.peer = model.dp[4].peer_group(int(d.n_params), dev)
        self.grad_loss = self.peer.grad_loss if self.peer is not None else \
            torch.zeros(int(d.n_params) + 4, dtype=torch.float32, device=dev)
        self.grad = self.grad_loss[:int(d.n_params)]
        self.loss = self.grad_loss[int(d.n_params):int(d.n_params) + 1]
        self.ws = torch.empty(max(16, int(model._loss_grad_workspace_bytes(batch_rows))), dtype=torch.uint8,
                              device=dev)
        self.perm = torch.empty(max(1, n_rows), dtype=torch.int32, device=dev)
        self.perm_ws = torch.empty(max(16, int(lib.flowmc_random_permutation_workspace_bytes(n_rows))),
                                   dtype=torch.uint8, device=dev)
        self.small = torch.empty(1024, dtype=torch.float32, device=dev)


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

    # ---- training (nf_model/base.py:98-210) -----------------------------------------------------
    # Data parallelism: ``dp = (rank, world_size, all_reduce[, broadcast])`` -- every rank holds the full training
    # set and the identical permutation, takes its contiguous slice of each global batch, and the flat
    # gradient + loss are sum-all-reduced before the (identical) optimiser step on every rank.  The all-reduce
    # result is bit-identical on every rank, so with identical initial parameters and identical whitening
    # constants (``train`` broadcasts data_mean / data_cov from rank 0) the replicas stay in lock-step.
    # FLOWMC_DP_MIN_ROWS (default 0 = always split): below this many rows per rank every rank takes the full
    # batch instead and the gradient is still all-reduced (averaged), so replicas cannot drift either way.
    dp = None
    dp_min_rows_per_rank = int(os.environ.get("FLOWMC_DP_MIN_ROWS", 0))

    # model-specific C-ABI calls behind loss_and_grad (the spline flow's; RealNVP overrides both)
    def _loss_grad_workspace_bytes(self, n_rows: int) -> int:
        return int(lib.flowmc_flow_loss_grad_workspace_bytes(C.byref(self.desc), int(n_rows)))

    def _loss_grad_call(self, x_ptr, idx_ptr, n, inv_n_total, grad_ptr, loss_ptr, ws_ptr, ws_bytes, stream):
        return lib.flowmc_flow_loss_grad(C.byref(self.desc), self.params.data_ptr(), x_ptr, idx_ptr, n, inv_n_total,
                                         grad_ptr, loss_ptr, ws_ptr, ws_bytes, stream)

    def prepare(self):
        """Hook: refresh derived device state after a parameter change (the spline flow's tensor-core image)."""

    def loss_and_grad(self, x, idx=None, scratch=None, n_global=None):
        """NFModel.loss_fn (base.py:98-100): (-mean log_prob, flat gradient).  ``idx`` (int32 device
        tensor) selects rows of ``x``."""
        x = x.contiguous()
        n = int(idx.numel()) if idx is not None else int(x.shape[0])
        sc = scratch or _TrainScratch(self, 0, n)
        self.prepare()     # refresh the tensor-core weight image: params change every step
        with torch.cuda.device(x.device):
            check(self._loss_grad_call(x.data_ptr(), idx.data_ptr() if idx is not None else None, n,
                                       1.0 / float(n_global or n), sc.grad.data_ptr(), sc.loss.data_ptr(),
                                       sc.ws.data_ptr(), sc.ws.numel(), _stream()))
        return sc.loss, sc.grad

    def _apply_update(self, optim, state, sc):
        state.count += 1
        check(lib.flowmc_clip_adamw(self.params.numel(), self.params.data_ptr(), sc.grad.data_ptr(),
                                    state.mu.data_ptr(), state.nu.data_ptr(), state.count,
                                    optim.learning_rate, optim.b1, optim.b2, optim.eps, optim.weight_decay,
                                    optim.max_norm, sc.small.data_ptr(), None, _stream()))

    def train_step(self, x, optim, state, idx=None, scratch=None):
        """One optimisation step IN PLACE on this model and ``state`` (base.py:102-125); returns the
        loss as a 1-element device tensor (no host synchronisation)."""
        n = int(idx.numel()) if idx is not None else int(x.shape[0])
        sc = scratch or _TrainScratch(self, 0, n)
        if self.dp is None or self.dp[1] == 1:
            self.loss_and_grad(x, idx, sc)
        elif n < self.dp[1] * self.dp_min_rows_per_rank:
            # opt-in (FLOWMC_DP_MIN_ROWS): every rank takes the whole batch, scaled 1 / world, and the all-reduce
            # averages the replicas' gradients -- the same update on every rank, bit for bit
            all_reduce = self.dp[2]
            self.loss_and_grad(x, idx, sc, n_global=n * self.dp[1])
            all_reduce(sc.grad_loss)
        else:
            rank, world, all_reduce = self.dp[:3]
            per = -(-n // world)
            lo, hi = min(n, rank * per), min(n, (rank + 1) * per)
            if idx is None:
                self.loss_and_grad(x[lo:hi], None, sc, n_global=n)
            else:
                self.loss_and_grad(x, idx[lo:hi], sc, n_global=n)
            if sc.peer is not None:
                # reduce-scatter + all-gather over NVLink peer memory, global norm, clip and AdamW in ONE kernel
                state.count += 1
                return sc.peer.step(self.params, state.mu, state.nu, state.count, optim, _stream())
            all_reduce(sc.grad_loss)      # gradient + loss, one collective
        self._apply_update(optim, state, sc)
        return sc.loss

    def _bind_train_step(self, x, optim, state, sc, n: int):
        """``train_step`` on ``n`` rows ``idx`` of ``x`` with everything that does not change between steps resolved
        once: returns ``f(idx_ptr) -> loss`` (device tensor; ``idx_ptr`` = device address of int32[n]).  Same calls,
        same order, same results as ``train_step``; the caller holds the device context."""
        x = x.contiguous()
        stream = _stream()
        x_ptr = x.data_ptr()
        dp = self.dp if (self.dp is not None and self.dp[1] > 1) else None
        lo, rows, n_glob = 0, n, n
        if dp is not None:
            if n < dp[1] * self.dp_min_rows_per_rank:
                n_glob = n * dp[1]
            else:
                per = -(-n // dp[1])
                lo = min(n, dp[0] * per)
                rows = min(n, (dp[0] + 1) * per) - lo
        inv = 1.0 / float(n_glob)
        grad_ptr, loss_ptr = sc.grad.data_ptr(), sc.loss.data_ptr()
        ws_ptr, ws_bytes = sc.ws.data_ptr(), sc.ws.numel()
        prepare = getattr(self, "_prepare_bound", None)
        prepare = prepare(stream) if prepare is not None else self.prepare
        loss_grad, apply_update = self._loss_grad_call, self._apply_update
        peer = sc.peer if dp is not None else None
        all_reduce = dp[2] if dp is not None else None
        keep = x                                       # the bound pointers stay valid while the closure lives

        def step(idx_ptr: int):
            prepare()
            check(loss_grad(x_ptr, idx_ptr + 4 * lo, rows, inv, grad_ptr, loss_ptr, ws_ptr, ws_bytes, stream))
            if peer is not None:
                state.count += 1
                return peer.step(self.params, state.mu, state.nu, state.count, optim, stream)
            if all_reduce is not None:
                all_reduce(sc.grad_loss)
            apply_update(optim, state, sc)
            return sc.loss

        step.keep = keep
        return step

    def train_epoch(self, rng, optim, state, data, batch_size, scratch=None):
        """base.py:127-151: permutation batches (incomplete tail skipped), in place; returns the last
        batch's loss (device tensor)."""
        n = int(data.shape[0])
        steps = n // int(batch_size)
        sc = scratch or _TrainScratch(self, n, int(batch_size) if steps > 0 else n)
        value = None
        if steps > 0:
            key = np.ascontiguousarray(rng, dtype=np.uint32)
            with torch.cuda.device(data.device):
                check(lib.flowmc_random_permutation(key.ctypes.data_as(_u32p), n, sc.perm.data_ptr(),
                                                    sc.perm_ws.data_ptr(), sc.perm_ws.numel(), _stream()))
            # every step of the epoch runs the same call sequence on a different slice of the permutation: bind it
            # once (pointers, stream, rank slice) -- the per-step host cost is what bounds a data-parallel rank whose
            # GPU work per step is a few hundred microseconds
            with torch.cuda.device(data.device):
                step = self._bind_train_step(data, optim, state, sc, int(batch_size))
                perm_ptr = sc.perm.data_ptr()
                for b in range(steps):
                    value = step(perm_ptr + 4 * b * int(batch_size))
        else:
            value = self.train_step(data, optim, state, None, sc)
        return value

    def train(self, rng, data, optim, state, num_epochs: int, batch_size: int, verbose: bool = True):
        """base.py:153-210.  Functional like the reference: ``self`` and ``state`` are left untouched;
        returns ``(rng, best_model, best_state, loss_values)`` where best = lowest last-batch loss."""
        data = torch.as_tensor(data, dtype=torch.float32)
        if not data.is_cuda:
            data = data.to(self.params.device)
        data = data.contiguous()
        n, d = data.shape
        batch_size = int(batch_size)
        steps = n // batch_size
        model = self.clone()
        model.dp = self.dp
        state = state.clone()
        best_model, best_state, best_loss = self, state.clone(), 1e9
        sc = _TrainScratch(model, n, batch_size if steps > 0 else n)
        with torch.cuda.device(data.device):                      # base.py:187-188
            check(lib.flowmc_data_mean_cov(data.data_ptr(), n, d, model.data_mean.data_ptr(),
                                           model.data_cov.data_ptr(), sc.small.data_ptr(), _stream()))
        if self.dp is not None and self.dp[1] > 1 and len(self.dp) > 3:
            # the moments are accumulated with float atomics (order-dependent last bits): rank 0's values are
            # THE whitening constants of every replica
            self.dp[3](model.data_mean)
            self.dp[3](model.data_cov)
        loss_values = np.zeros(num_epochs, np.float32)
        rng = np.asarray(rng, dtype=np.uint32)
        it = range(num_epochs)
        if verbose:
            from tqdm import trange
            it = trange(num_epochs, desc="Training NF", miniters=max(1, int(num_epochs / 10)))
        for epoch in it:
            rng, input_rng = frandom.split(rng)
            with nvtx_range(f"flowmc/train_epoch[{epoch}]"):
                value = model.train_epoch(input_rng, optim, state, data, batch_size, sc)
            loss_values[epoch] = float(value.item())              # the reference's per-epoch host read
            if sc.peer is not None and sc.peer.failed():
                # a cross-GPU barrier of flowmc_dp_reduce_adamw gave up waiting (bounded spin): the parameters of
                # this epoch are not trustworthy -- fail loudly instead of sampling from them
                raise RuntimeError("flowmc_b200: a peer rank did not reach the data-parallel optimiser step "
                                   "(flowmc_dp_reduce_adamw barrier timeout)")
            if loss_values[epoch] < best_loss:
                if best_model is self:
                    best_model = model.clone()
                else:
                    best_model.params.copy_(model.params)
                best_state.copy_(state)
                best_loss = loss_values[epoch]
        best_model.dp = self.dp
        return rng, best_model, best_state, torch.from_numpy(loss_values).to(data.device)

    def to_precision(self, precision: str = "float32"):
        """The B200 path computes in float32 only (nf_model/base.py:212-242 is experimental upstream)."""
        if precision.lower() != "float32":
            raise NotImplementedError("flowmc_b200 flows are float32")
        return self
