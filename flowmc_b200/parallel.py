"""Multi-GPU plumbing: one process per GPU, chains sharded with no communication during local and
global steps, flow training data-parallel.

The reference has no multi-device code at all; this is the B200 addition that
BASELINE.json's north_star asks for:
  * chains [offset, offset + n_local) live on this rank; per-chain PRNG keys are taken from the
    GLOBAL ``split(subkey, n_chains)`` so any sharding reproduces the single-GPU chains bit for bit;
  * TrainModel: every rank gathers the selected training rows that belong to its chains into a
    zero-filled [n_max_examples, d] buffer and a sum-all-reduce assembles the full set on every rank
    (an all-gather of disjoint rows); training then splits each global batch across ranks and
    sum-all-reduces the flat gradient vector (NCCL over NVLink; gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class ChainShard:
    def __init__(self, n_chains_global: int, rank: int | None = None, world_size: int | None = None, group=None):
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else int(rank)
        self.world_size = dist.get_world_size(group) if world_size is None else int(world_size)
        self.n_chains_global = int(n_chains_global)
        per = -(-self.n_chains_global // self.world_size)
        self.offset = min(self.n_chains_global, self.rank * per)
        self.n_local = min(self.n_chains_global, self.offset + per) - self.offset

    def slab(self, x: torch.Tensor) -> torch.Tensor:
        """This rank's rows of a [n_chains_global, ...] tensor."""
        return x[self.offset:self.offset + self.n_local]

    def all_reduce(self, t: torch.Tensor, op: str = "sum"):
        if self.world_size > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM, group=self.group)
        return t

    def attach(self, local_stepper, global_stepper, model_trainer, model):
        local_stepper.set_chain_shard(self.offset, self.n_chains_global)
        global_stepper.set_chain_shard(self.offset, self.n_chains_global)
        model_trainer.set_chain_shard(self.offset, self.n_chains_global, self.all_reduce)
        model.dp = (self.rank, self.world_size, self.all_reduce)

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
