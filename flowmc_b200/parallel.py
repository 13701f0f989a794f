"""Multi-GPU plumbing: one process per GPU, chains sharded with no communication during local and
global steps, flow training data-parallel.

The reference has no multi-device code at all; this is the B200 addition that
BASELINE.json's north_star asks for:
  * chains [offset, offset + n_local) live on this rank; per-chain PRNG keys are taken from the
    GLOBAL ``split(subkey, n_chains)`` so any sharding reproduces the single-GPU chains bit for bit;
  * TrainModel: every rank gathers the selected training rows that belong to its chains into its block of a
    [world, rows_per_rank, d] buffer and an all-gather (``all_gather_into_tensor``) gives every rank all blocks,
    which are then put back into the order of the global ``choice`` draw; training splits each global batch
    across ranks and sum-all-reduces the flat gradient vector + loss (NCCL over NVLink; gloo in the CPU tests);
    ``data_mean`` / ``data_cov`` are broadcast from rank 0 so that every replica whitens identically.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.distributed as dist


class _DevMem:
    """Device memory that torch did not allocate, exposed through __cuda_array_interface__ (float32 vector)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class PeerGroup:
    """The ranks' exchange blocks for ``flowmc_dp_reduce_adamw`` (csrc/peer_reduce.cu): one CUDA-IPC allocation per
    rank, mapped into every peer over NVLink.  Collective constructor (handles travel through torch.distributed)."""

    def __init__(self, rank: int, world: int, n_params: int, device, group=None):
        from ._lib import check, lib
        self.rank, self.world, self.n_params = int(rank), int(world), int(n_params)
        nbytes = int(lib.flowmc_peer_block_bytes(self.n_params))
        if nbytes <= 0:
            raise ValueError("peer exchange needs n_params % 4 == 0")
        mine = C.c_void_p()
        handle = C.create_string_buffer(64)
        with torch.cuda.device(device):
            check(lib.flowmc_ipc_alloc(nbytes, C.byref(mine), handle))
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle.raw), group=group)
            self.blocks = (C.c_void_p * self.world)()
            self._opened = []
            for k in range(self.world):
                if k == self.rank:
                    self.blocks[k] = mine.value
                else:
                    p = C.c_void_p()
                    check(lib.flowmc_ipc_open(handles[k], C.byref(p)))
                    self.blocks[k] = p.value
                    self._opened.append(p.value)
        self._mine = mine.value
        self.epoch = 0
        off = int(lib.flowmc_peer_block_offset(self.n_params, 0))
        # the backward pass writes this rank's gradient (+ loss at index n_params) straight into the exchange block
        self._mem = _DevMem(self._mine + off, self.n_params + 4)
        self.grad_loss = torch.as_tensor(self._mem, device=device)
        self.loss_out = torch.zeros(1, dtype=torch.float32, device=device)
        dist.barrier(group=group)      # every peer has opened every block before anyone launches on them

    def step(self, params, mu, nu, count, optim, stream) -> torch.Tensor:
        """All-reduce of the ranks' gradients + clip + AdamW in one kernel; returns the all-reduced loss (device)."""
        from ._lib import check, lib
        self.epoch += 1
        check(lib.flowmc_dp_reduce_adamw(self.rank, self.world, self.blocks, self.n_params, params.data_ptr(),
                                         mu.data_ptr(), nu.data_ptr(), int(count), optim.learning_rate, optim.b1,
                                         optim.b2, optim.eps, optim.weight_decay, optim.max_norm, self.epoch,
                                         self.loss_out.data_ptr(), stream))
        return self.loss_out

    def failed(self) -> bool:
        """True if a barrier of the kernel gave up waiting for a peer (bounded spin)."""
        from ._lib import lib
        off = int(lib.flowmc_peer_block_offset(self.n_params, 4))
        flag = torch.as_tensor(_DevMem(self._mine + off, 4), device=self.grad_loss.device)
        return bool(flag.view(torch.int32)[3].item() != 0)


class ChainShard:
    def __init__(self, n_chains_global: int, rank: int | None = None, world_size: int | None = None, group=None):
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else int(rank)
        self.world_size = dist.get_world_size(group) if world_size is None else int(world_size)
        self.n_chains_global = int(n_chains_global)
        per = -(-self.n_chains_global // self.world_size)
        self.offset = min(self.n_chains_global, self.rank * per)
        self.n_local = min(self.n_chains_global, self.offset + per) - self.offset

    def _staged(self, t: torch.Tensor) -> bool:
        """gloo moves host memory: CUDA tensors are staged through the host (single-GPU multi-process tests)."""
        return t.is_cuda and dist.get_backend(self.group) == "gloo"

    def slab(self, x: torch.Tensor) -> torch.Tensor:
        """This rank's rows of a [n_chains_global, ...] tensor."""
        return x[self.offset:self.offset + self.n_local]

    def all_reduce(self, t: torch.Tensor, op: str = "sum"):
        if self.world_size > 1:
            rop = dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM
            if self._staged(t):
                h = t.cpu()
                dist.all_reduce(h, op=rop, group=self.group)
                t.copy_(h)
            else:
                dist.all_reduce(t, op=rop, group=self.group)
        return t

    def broadcast(self, t: torch.Tensor, src: int = 0):
        if self.world_size > 1:
            gsrc = dist.get_global_rank(self.group, src) if self.group is not None else src
            if self._staged(t):
                h = t.cpu()
                dist.broadcast(h, src=gsrc, group=self.group)
                t.copy_(h)
            else:
                dist.broadcast(t, src=gsrc, group=self.group)
        return t

    def all_gather_blocks(self, block: torch.Tensor) -> torch.Tensor:
        """Every rank contributes one equally-shaped ``block``; returns [world_size, *block.shape]."""
        if self.world_size == 1:
            return block.unsqueeze(0)
        src = block.cpu() if self._staged(block) else block.contiguous()
        out = torch.empty((self.world_size * block.shape[0],) + tuple(block.shape[1:]), dtype=block.dtype,
                          device=src.device)
        dist.all_gather_into_tensor(out, src, group=self.group)
        return out.to(block.device).view((self.world_size,) + tuple(block.shape))

    def owner_of_chain(self, chain: torch.Tensor) -> torch.Tensor:
        per = -(-self.n_chains_global // self.world_size)
        return torch.div(chain, per, rounding_mode="floor")

    def assemble_rows(self, idx: torch.Tensor, window: int, d: int, gather_own) -> torch.Tensor:
        """All-gather assembly of the training set.  ``idx`` [m] is the global ``choice`` draw, known to
        every rank (row q of the population belongs to global chain q // window).  Rows are grouped by owning rank
        (stable), ``gather_own(my_idx, block)`` fills this rank's block with the rows of its own chains, one
        ``all_gather_into_tensor`` moves every block to every rank, and the blocks go back into draw order."""
        m = int(idx.numel())
        owner = self.owner_of_chain(torch.div(idx, window, rounding_mode="floor")).to(torch.int64)
        order = torch.argsort(owner, stable=True)
        counts = [int(c) for c in torch.bincount(owner, minlength=self.world_size).tolist()]
        starts = [0]
        for c in counts:
            starts.append(starts[-1] + c)
        cap = max(1, max(counts))
        mine = order[starts[self.rank]:starts[self.rank + 1]]
        block = torch.zeros((cap, d), dtype=torch.float32, device=idx.device)
        if mine.numel():
            gather_own(idx[mine].contiguous(), block)
        blocks = self.all_gather_blocks(block)                          # [world, cap, d]
        out = torch.empty((m, d), dtype=torch.float32, device=idx.device)
        for r in range(self.world_size):
            if counts[r]:
                out[order[starts[r]:starts[r + 1]]] = blocks[r, :counts[r]]
        return out

    def peer_group(self, n_params: int, device):
        """The NVLink exchange group for the fused data-parallel optimiser step, or None when it does not apply: it
        needs one process per GPU of ONE box over NCCL (world <= 8); ``FLOWMC_DP_PEER=0`` keeps NCCL + flowmc_clip_adamw."""
        if (self.world_size < 2 or self.world_size > 8 or os.environ.get("FLOWMC_DP_PEER", "1") == "0"
                or dist.get_backend(self.group) != "nccl" or torch.cuda.device_count() < self.world_size
                or n_params % 4 != 0):
            return None
        key = (int(n_params), str(device))
        if not hasattr(self, "_peer_groups"):
            self._peer_groups = {}
        if key not in self._peer_groups:
            self._peer_groups[key] = PeerGroup(self.rank, self.world_size, n_params, device, self.group)
        return self._peer_groups[key]

    def attach(self, local_stepper, global_stepper, model_trainer, model):
        local_stepper.set_chain_shard(self.offset, self.n_chains_global)
        global_stepper.set_chain_shard(self.offset, self.n_chains_global)
        model_trainer.set_chain_shard(self.offset, self.n_chains_global, self.all_reduce, self)
        model.dp = (self.rank, self.world_size, self.all_reduce, self.broadcast, self)

    def gather_chains(self, x: torch.Tensor) -> torch.Tensor:
        """All ranks' slabs concatenated along the chain axis (for users who want the full buffer)."""
        if self.world_size == 1:
            return x
        per = -(-self.n_chains_global // self.world_size)
        pad = torch.zeros((per,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        pad[:x.shape[0]] = x
        out = [torch.empty_like(pad) for _ in range(self.world_size)]
        dist.all_gather(out, pad, group=self.group)
        return torch.cat(out)[:self.n_chains_global]
