"""Full flowMC Sampler (RQSpline_MALA_Bundle) on the B200 path: BASELINE.json configs[4] (C5) by default.

  python scripts/bench_sampler.py [--chains 65536] [--dim 64] [--quick]
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_sampler.py   # chains sharded

C5: 64-D mixture of 8 Gaussians, 65536 chains, n_local_steps=50, n_global_steps=10, 4 training + 4
production loops, n_epochs=5, batch_size=16384, n_max_examples=1,048,576, flow 8x[128,128]x8 bins.
Prints one JSON line (rank 0): wall time per phase (CUDA events around each strategy), chain-steps/s, flow-train
samples/s and a Geyer ESS/s estimate on the production samples.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def geyer_ess(x):
    """Initial-positive-sequence ESS of chains x[c, t] (one dimension), summed over chains (numpy)."""
    c, t = x.shape
    x = x - x.mean(axis=1, keepdims=True)
    f = np.fft.rfft(x, n=2 * t, axis=1)
    acov = np.fft.irfft(f * np.conj(f), axis=1)[:, :t] / t
    rho = acov / np.maximum(acov[:, :1], 1e-30)
    pairs = rho[:, 0:t - 1:2] + rho[:, 1:t:2]
    pos = np.cumprod(pairs > 0, axis=1)
    tau = -1.0 + 2.0 * np.sum(pairs * pos, axis=1)
    return float(np.sum(t / np.maximum(tau, 1.0)))


class Timed:
    def __init__(self, name, inner, acc):
        self.name, self.inner, self.acc = name, inner, acc

    def __call__(self, *a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = self.inner(*a)
        e1.record()
        self.acc.setdefault(self.name, []).append((e0, e1))
        return out


def run_sampler(dev, rank, world, n_chains=65536, d=64, quick=False, flow=(8, [128, 128])):
    """One full C5 run (strong scaling: ``n_chains`` in TOTAL, sharded over ``world`` ranks).  Collective: every rank
    must call it.  Returns the result dict (identical on every rank apart from rank-0 phase times)."""
    from flowmc_b200 import random as frandom, targets as T
    from flowmc_b200.parallel import ChainShard
    from flowmc_b200.resource_strategy_bundle.RQSpline_MALA import RQSpline_MALA_Bundle
    from flowmc_b200.Sampler import Sampler

    cfg = dict(n_local_steps=50, n_global_steps=10, n_training_loops=4, n_production_loops=4, n_epochs=5,
               mala_step_size=0.1, rq_spline_hidden_units=list(flow[1]), rq_spline_n_bins=8, rq_spline_n_layers=int(flow[0]),
               learning_rate=1e-3, batch_size=16384, n_max_examples=1048576)
    if quick:
        cfg.update(n_training_loops=2, n_production_loops=2, n_epochs=2, n_max_examples=131072)
    rs = np.random.RandomState(0)
    mu = np.zeros((8, d), np.float32)
    for i, ax in enumerate(rs.choice(d, 8, replace=False)):
        mu[i, ax] = 3.0 if i % 2 == 0 else -3.0
    target = T.gaussian_mixture(mu, 1.0)

    shard = ChainShard(n_chains, rank, world) if world > 1 else None
    key = frandom.PRNGKey(42)
    key, sub = frandom.split(key)
    x0 = frandom.normal(sub, (n_chains, d), device=dev)
    if shard is not None:
        x0 = shard.slab(x0).contiguous()
    # warm-up: a miniature run of the same bundle (same target, dimension and flow shape, 256 chains per rank) so that
    # CUDA's lazy module loading (~50 ms for the first launch of each kernel) and the NCCL communicators are not timed
    wcfg = dict(cfg, n_training_loops=1, n_production_loops=1, n_epochs=1, batch_size=128 * world,
                n_max_examples=256 * world)
    wshard = ChainShard(256 * world, rank, world) if world > 1 else None
    wx0 = frandom.normal(frandom.PRNGKey(1), (256 * world, d), device=dev)
    wbundle = RQSpline_MALA_Bundle(frandom.PRNGKey(2), 256 * world, d, target, chain_shard=wshard, **wcfg)
    Sampler(d, 256 * world, frandom.PRNGKey(3), resource_strategy_bundles=wbundle).sample(
        wshard.slab(wx0).contiguous() if wshard is not None else wx0, {})
    del wbundle
    key, sub = frandom.split(key)
    bundle = RQSpline_MALA_Bundle(sub, n_chains, d, target, chain_shard=shard, **cfg)
    acc = {}
    for nm in ("local_stepper", "global_stepper", "model_trainer"):
        bundle.strategies[nm] = Timed(nm, bundle.strategies[nm], acc)
    sampler = Sampler(d, n_chains, key, resource_strategy_bundles=bundle)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    sampler.sample(x0, {})
    ev1.record()
    torch.cuda.synchronize()
    wall_host = time.perf_counter() - t0
    tt = torch.tensor([ev0.elapsed_time(ev1) * 1e-3], device=dev, dtype=torch.float64)   # device time, max over ranks
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.barrier()
    wall = float(tt.item())
    ms = {k: sum(a.elapsed_time(b) for a, b in v) for k, v in acc.items()}
    res = sampler.resources
    loops = cfg["n_training_loops"] + cfg["n_production_loops"]
    local_steps = n_chains * cfg["n_local_steps"] * loops
    global_steps = n_chains * cfg["n_global_steps"] * loops
    train_steps = cfg["n_training_loops"] * cfg["n_epochs"] * (cfg["n_max_examples"] // cfg["batch_size"])
    train_rows = train_steps * cfg["batch_size"]
    ga = res["global_accs_production"].data
    la = res["local_accs_production"].data
    pos = res["positions_production"].data
    n_ess = min(pos.shape[0], 256)
    ess = min(geyer_ess(pos[:n_ess, :, j].cpu().numpy()) for j in range(0, d, max(1, d // 8))) * pos.shape[0] / n_ess
    if world > 1:
        t = torch.tensor([ess], device=dev, dtype=torch.float64)
        dist.all_reduce(t)
        ess = float(t.item())
    return {
        "workload": f"C5: full Sampler (MALA + NFProposal + TrainModel), {d}-D 8-component mixture, {n_chains} chains "
                    f"in total over {world} GPU(s) (strong scaling), flow {flow[0]}x{list(flow[1])}x8",
        "config": cfg, "n_gpus": world, "wall_s": wall, "wall_s_host_rank0": wall_host, "phase_ms_rank0": ms,
        "timing": "CUDA events around Sampler.sample on each rank, max over ranks",
        "local_chain_steps_per_s": local_steps / (ms["local_stepper"] * 1e-3),
        "global_chain_steps_per_s": global_steps / (ms["global_stepper"] * 1e-3),
        "flow_train_samples_per_s": train_rows / (ms["model_trainer"] * 1e-3),
        "flow_train_ms_per_step": ms["model_trainer"] / train_steps,
        "flow_train_parallelism": ("data-parallel: batch split over ranks, gradient + loss all-reduce (NCCL)"
                                   if world > 1 else "single GPU"),
        "sampler_chain_steps_per_s": (local_steps + global_steps) / wall,
        "ess_per_s": ess / wall,
        "global_acceptance": float(ga[torch.isfinite(ga)].mean()), "local_acceptance": float(la[torch.isfinite(la)].mean()),
        "final_loss": float(res["loss_buffer"].data[-1]), "first_loss": float(res["loss_buffer"].data[0]),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=65536)
    ap.add_argument("--dim", type=int, default=64)
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    out = run_sampler(dev, rank, world, args.chains, args.dim, args.quick)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
