"""The product's multi-rank branches on the GPU with world_size 2: data-parallel ``NFModel.train``
(gradient + loss all-reduce, data_mean / data_cov broadcast), ``TrainModel``'s all-gather assembly of the training
set, sharded local / global steps inside a full bundle run, and ParallelTempering's cross-rank ladder adaptation.

Two processes are spawned: over NCCL when the box has >= 2 GPUs (gpurun --gpus 2), otherwise both on cuda:0 over
gloo with host staging (flowmc_b200.parallel.ChainShard stages CUDA tensors for gloo), so the branches run in the
single-GPU round-end test tier as well.  Reference sides: strategy/train_model.py:47-112,
resource/model/nf_model/base.py:127-210 (single device there; the split is the B200 addition)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

# data-parallel vs single-process parameters after training: the two sum the same per-row gradients in a different
# order (two half-batch partial sums instead of one), ~1e-7 relative per step, fed back through AdamW's normalised
# update for 8 steps
DP_PARAM_RTOL = 5e-5


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    nccl = torch.cuda.device_count() >= world
    dev = torch.device("cuda", rank if nccl else 0)
    torch.cuda.set_device(dev)
    if nccl:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from flowmc_b200 import random as frandom, targets as T
        from flowmc_b200.parallel import ChainShard
        from flowmc_b200.resource.model.nf_model.rqSpline import MaskedCouplingRQSpline
        from flowmc_b200.resource.optimizer import Optimizer
        from flowmc_b200.resource_strategy_bundle.RQSpline_MALA import RQSpline_MALA_Bundle
        from flowmc_b200.Sampler import Sampler
        out = {"backend": "nccl" if nccl else "gloo(staged)"}

        # ---- 1. data-parallel NFModel.train == single-process train -----------------------------------------
        d, n_rows, bs = 8, 4096, 1024
        shard = ChainShard(64, rank, world)
        data = frandom.normal(frandom.PRNGKey(5), (n_rows, d), device=dev) * 1.5 + 0.3
        m1 = MaskedCouplingRQSpline(d, 3, [32, 32], 8, frandom.PRNGKey(1), device=dev)
        o1 = Optimizer(m1, 2e-3)
        _, best1, st1, loss1 = m1.train(frandom.PRNGKey(2), data, o1.optim, o1.optim_state, 2, bs, verbose=False)
        m2 = MaskedCouplingRQSpline(d, 3, [32, 32], 8, frandom.PRNGKey(1), device=dev)
        m2.dp = (shard.rank, shard.world_size, shard.all_reduce, shard.broadcast)
        o2 = Optimizer(m2, 2e-3)
        _, best2, st2, loss2 = m2.train(frandom.PRNGKey(2), data, o2.optim, o2.optim_state, 2, bs, verbose=False)
        scale = float(best1.params.abs().max())
        err = float((best1.params - best2.params).abs().max())
        assert err <= DP_PARAM_RTOL * scale, f"DP params differ from the single-process run by {err:.3e} (scale {scale:.3g})"
        assert float((loss1 - loss2).abs().max()) <= 1e-5 * max(1.0, float(loss1.abs().max()))
        # replicas are bit-identical to each other (same all-reduced gradient, same whitening constants)
        both = shard.all_gather_blocks(best2.params)
        assert torch.equal(both[0], both[1]), "data-parallel replicas drifted apart"
        mu = shard.all_gather_blocks(st2.mu)
        assert torch.equal(mu[0], mu[1])
        out["dp_param_err"] = err / scale

        # ---- 1b. the same training with the collective INSIDE the optimiser kernel (NVLink peer memory, CUDA IPC):
        #          flowmc_dp_reduce_adamw instead of NCCL all-reduce + flowmc_clip_adamw.  Needs one GPU per rank.
        if nccl:
            m3 = MaskedCouplingRQSpline(d, 3, [32, 32], 8, frandom.PRNGKey(1), device=dev)
            m3.dp = (shard.rank, shard.world_size, shard.all_reduce, shard.broadcast, shard)
            o3 = Optimizer(m3, 2e-3)
            _, best3, st3, loss3 = m3.train(frandom.PRNGKey(2), data, o3.optim, o3.optim_state, 2, bs, verbose=False)
            pg = shard.peer_group(int(m3.desc.n_params), dev)
            assert pg is not None and pg.epoch == 2 * (n_rows // bs) and not pg.failed(), "peer path not taken"
            err3 = float((best1.params - best3.params).abs().max())
            assert err3 <= DP_PARAM_RTOL * scale, f"peer-memory DP params differ by {err3:.3e} (scale {scale:.3g})"
            assert float((loss1 - loss3).abs().max()) <= 1e-5 * max(1.0, float(loss1.abs().max()))
            both = shard.all_gather_blocks(best3.params)
            assert torch.equal(both[0], both[1]), "peer-memory replicas drifted apart"
            for t3 in (st3.mu, st3.nu):
                tt = shard.all_gather_blocks(t3)
                assert torch.equal(tt[0], tt[1])
            assert st3.count == st1.count
            out["peer_param_err"] = err3 / scale

        # ---- 2. full bundle, chains sharded: buffers == rows of the single-process run until training, and the
        #         all-gathered training set == the single-process selection ------------------------------------
        n_chains, dd = 64, 5
        cfg = dict(n_local_steps=12, n_global_steps=4, n_training_loops=2, n_production_loops=1, n_epochs=2,
                   mala_step_size=0.1, rq_spline_hidden_units=[32, 32], rq_spline_n_bins=8, rq_spline_n_layers=3,
                   learning_rate=1e-3, batch_size=256, n_max_examples=1024)
        target = T.dual_moon()
        x0 = frandom.normal(frandom.PRNGKey(9), (n_chains, dd), device=dev)
        full = RQSpline_MALA_Bundle(frandom.PRNGKey(3), n_chains, dd, target, **cfg)
        Sampler(dd, n_chains, frandom.PRNGKey(4), resource_strategy_bundles=full).sample(x0, {})
        sh = ChainShard(n_chains, rank, world)
        part = RQSpline_MALA_Bundle(frandom.PRNGKey(3), n_chains, dd, target, chain_shard=sh, **cfg)
        Sampler(dd, n_chains, frandom.PRNGKey(4), resource_strategy_bundles=part).sample(sh.slab(x0).contiguous(), {})
        pf = full.resources["positions_training"].data
        pp = part.resources["positions_training"].data
        first_loop = cfg["n_local_steps"]          # local steps of loop 1 precede any training: bit-identical
        assert torch.equal(pp[:, :first_loop], sh.slab(pf)[:, :first_loop])
        tf = full.strategies["model_trainer"].last_training_data
        tp = part.strategies["model_trainer"].last_training_data
        assert tf.shape == tp.shape == (cfg["n_max_examples"], dd)
        # the LAST TrainModel call sees buffers produced after a trained flow (DP and single-process flows differ in the
        # last bits, which can flip a global accept): the assembly is compared bit-exactly on identical buffers instead
        # (not the sampler's own buffer: data_mean / data_cov are accumulated with float atomics, so two PROCESSES
        # running the same single-GPU sampler agree only to the last bits, and a flipped global accept changes rows)
        pf = frandom.normal(frandom.PRNGKey(21), (n_chains, 24, dd), device=dev)
        pf[:, 20:] = float("-inf")                  # unfilled tail of the buffer: skipped by the finite-row filter
        trainer = part.strategies["model_trainer"]
        saved = (trainer.chain_shard, trainer.shard)
        _, sel_part = trainer.select_training_data(frandom.PRNGKey(11), sh.slab(pf).contiguous())
        trainer.chain_shard, trainer.shard = None, None
        _, sel_full = trainer.select_training_data(frandom.PRNGKey(11), pf)
        trainer.chain_shard, trainer.shard = saved
        assert torch.equal(sel_part, sel_full), "all-gather assembly differs from the single-process selection"
        # legacy assembly (sum-all-reduce of a zero-filled buffer) gives the same rows
        trainer.shard = None
        _, sel_sum = trainer.select_training_data(frandom.PRNGKey(11), sh.slab(pf).contiguous())
        trainer.shard = saved[1]
        assert torch.equal(sel_sum, sel_full)
        fl = full.resources["loss_buffer"].data
        pl = part.resources["loss_buffer"].data
        ne = cfg["n_epochs"]                          # loop 1 trains on bit-identical data
        assert torch.isfinite(pl).all()
        assert float((fl[:ne] - pl[:ne]).abs().max()) <= 1e-4 * max(1.0, float(fl[:ne].abs().max()))
        assert float((fl - pl).abs().max()) <= 5e-2 * max(1.0, float(fl.abs().max()))
        ret[rank] = out
    except Exception:  # pragma: no cover
        import traceback
        ret[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


def test_world_size_2_product_branches(cuda):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        res = dict(ret)
        assert all(isinstance(res.get(r), dict) for r in range(world)), res
        print(res)
