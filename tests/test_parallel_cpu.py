"""Multi-process host logic on CPU (gloo, world_size 2): chain partitioning, the all-reduce-as-all-gather
assembly of the training set, and data-parallel gradient identity -- with the oracle standing in for the
per-rank device work (the CUDA kernels themselves are covered by the -m gpu tests, including bit-identical
chain sharding in tests/test_gpu_nf.py and tests/test_gpu_local.py)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from flowmc_b200.parallel import ChainShard
        from oracle import flow as oflow, local as olocal, nf, rng, targets as otargets
        n_chains, d, n_steps = 10, 4, 6
        shard = ChainShard(n_chains)
        assert shard.world_size == world and shard.n_local == 5 and shard.offset == 5 * rank

        # 1. sharded local steps == the same chains of the unsharded run (global key indexing, no comms)
        key = rng.PRNGKey(3)
        x0 = rng.normal(rng.PRNGKey(4), (n_chains, d))
        packed = otargets.IsoGaussian.pack(d, 0.5)
        k = olocal.make_kernel("MALA", step_size=0.3)
        full = olocal.take_serial_steps(key, x0, "iso_gaussian", packed, k, n_steps)
        mine = olocal.take_serial_steps(key, shard.slab(torch.from_numpy(x0)).numpy(), "iso_gaussian", packed, k,
                                        n_steps, chain_offset=shard.offset, n_chains_total=n_chains)
        assert np.array_equal(mine[1], full[1][shard.offset:shard.offset + shard.n_local])
        assert np.array_equal(mine[0], full[0])
        gathered = shard.gather_chains(torch.from_numpy(mine[1]))
        assert np.array_equal(gathered.numpy(), full[1])

        # 2. training-set assembly: every rank fills the rows of its own chains, sum-all-reduce = all-gather
        buf = full[1]                                              # [n_chains, n_steps, d]
        tkey = rng.PRNGKey(5)
        _, _, want, idx = nf.select_training_data(tkey, buf, 64, 4)
        window = 4
        out = torch.zeros((64, d))
        chain_of = idx // window
        own = (chain_of >= shard.offset) & (chain_of < shard.offset + shard.n_local)
        rows = buf[:, -window:].reshape(-1, d)[idx]
        out[torch.from_numpy(own)] = torch.from_numpy(rows[own])
        shard.all_reduce(out)
        assert np.array_equal(out.numpy(), want)

        # 2b. the product's all-gather assembly (ChainShard.assemble_rows) with a numpy gather standing in for
        #     flowmc_gather_training_rows: same rows in the same order as the single-process selection
        pop = buf[:, -window:].reshape(-1, d)

        def gather_own(my_idx, block):
            ch = my_idx.numpy() // window
            assert ((ch >= shard.offset) & (ch < shard.offset + shard.n_local)).all()
            block[:my_idx.numel()] = torch.from_numpy(pop[my_idx.numpy()])
        got = shard.assemble_rows(torch.from_numpy(idx.astype(np.int32)), window, d, gather_own)
        assert np.array_equal(got.numpy(), want)

        # 2c. ParallelTempering._adapt_temperature: per-rung accept counts summed over ranks -> the ladder of the
        #     unsharded run on every rank (parallel_tempering.py:400-436)
        from flowmc_b200.strategy.parallel_tempering import ParallelTempering
        rs = np.random.RandomState(0)
        acc_full = (rs.rand(n_chains, 4) < np.array([0.2, 0.5, 0.8, 0.4])).astype(np.float32)
        temps = torch.tensor([1.0, 2.0, 4.0, 8.0, 16.0])
        pt1 = ParallelTempering(n_steps=1, tempered_logpdf_name="t", kernel_name="k",
                                tempered_buffer_names=["a", "b"], state_name="s")
        want_t = pt1._adapt_temperature(temps, torch.from_numpy(acc_full))
        ptr = ParallelTempering(n_steps=1, tempered_logpdf_name="t", kernel_name="k",
                                tempered_buffer_names=["a", "b"], state_name="s")
        ptr.set_chain_shard(shard.offset, n_chains, shard.all_reduce)
        got_t = ptr._adapt_temperature(temps, shard.slab(torch.from_numpy(acc_full)))
        assert torch.equal(got_t, want_t), (got_t, want_t)
        assert not torch.equal(want_t, temps)

        # 2d. broadcast keeps replicas' whitening constants identical
        mom = torch.full((3,), float(rank + 1))
        shard.broadcast(mom)
        assert torch.equal(mom, torch.ones(3))

        # 3. data-parallel gradient: per-rank slices scaled by 1/global batch, summed == full-batch gradient
        p = oflow.init_params(rng.PRNGKey(1), d, 2, [8, 8], 4)
        x = want[:32]
        loss_full, g_full = nf.loss_and_grads(p, x)
        per = 16
        xs = x[rank * per:(rank + 1) * per]
        loss_r, g_r = nf.loss_and_grads(p, xs)
        flat = torch.from_numpy(nf.flatten(g_r, p).astype(np.float64) * (per / 32))
        loss_t = torch.tensor([loss_r * per / 32], dtype=torch.float64)
        shard.all_reduce(flat)
        shard.all_reduce(loss_t)
        np.testing.assert_allclose(flat.numpy(), nf.flatten(g_full, p), rtol=1e-5, atol=1e-8)
        assert abs(loss_t.item() - loss_full) < 1e-6 * max(1.0, abs(loss_full))
        mm = torch.tensor([float(rank)])
        shard.all_reduce(mm, "max")
        assert mm.item() == world - 1
        ret[rank] = "ok"
    except Exception as ex:  # pragma: no cover
        import traceback
        ret[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: "ok", 1: "ok"}, dict(ret)


def test_chain_shard_partition_covers_all_chains():
    from flowmc_b200.parallel import ChainShard
    for n, w in ((65536, 8), (10, 4), (7, 8), (1, 2)):
        spans = [(s.offset, s.n_local) for s in (ChainShard(n, r, w) for r in range(w))]
        assert sum(c for _, c in spans) == n
        pos = 0
        for off, c in spans:
            assert off == pos or c == 0
            pos += c
