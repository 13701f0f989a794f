"""Edge cases of the hot path: empty and single-row batches, tile boundaries of the 128-row tensor-core tiles,
one chain / one dimension / one step for the local kernels, a one-step global call."""
import numpy as np
import pytest
import torch

from flowutil import model_from_params, random_params
from parity import assert_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("path", ["0", "3"])
@pytest.mark.parametrize("n", [127, 128, 129, 257])
def test_flow_tile_boundaries(cuda, monkeypatch, path, n):
    """Row counts around the 128-row tile of the tensor-core path (ragged last tile, exactly full tiles)."""
    from oracle import flow as oflow
    monkeypatch.setenv("FLOWMC_FLOW_TC", path)
    p = random_params(11, 32, 2, [128, 128], 8)
    m = model_from_params(p)
    r = np.random.default_rng(n)
    x = (2.0 * r.standard_normal((n, 32))).astype(np.float32)
    lp = m.log_prob(torch.from_numpy(x).cuda())
    o32 = oflow.log_prob(p, x)
    with oflow.precision(np.float64):
        o64 = oflow.log_prob(p, x)
    assert lp.shape == (n,)
    assert_close(lp.cpu().numpy(), o32, "log_prob", floor=o32 - o64, floor_factor=12.0 if path == "3" else 3.0)
    # rows are independent: the last (ragged) tile's rows equal the same rows evaluated on their own
    tail = m.log_prob(torch.from_numpy(x[-5:]).cuda())
    assert torch.allclose(tail, lp[-5:], rtol=0, atol=0)


@pytest.mark.parametrize("path", ["0", "3"])
def test_flow_empty_batch(cuda, monkeypatch, path):
    from oracle import rng
    monkeypatch.setenv("FLOWMC_FLOW_TC", path)
    m = model_from_params(random_params(2, 5, 2, [32, 32], 8))
    x = torch.empty((0, 5), device="cuda")
    assert m.log_prob(x).shape == (0,)
    y, ld = m.forward(x)
    assert y.shape == (0, 5) and ld.shape == (0,)
    assert m.sample(rng.PRNGKey(1), 0).shape == (0, 5)


@pytest.mark.parametrize("kind", ["MALA", "GRW", "HMC"])
@pytest.mark.parametrize("n_chains,d,n_steps", [(1, 1, 1), (1, 3, 7), (3, 1, 33)])
def test_local_kernels_smallest_shapes(cuda, kind, n_chains, d, n_steps):
    from flowmc_b200 import random as frandom, targets as T
    from flowmc_b200.resource.kernel.Gaussian_random_walk import GaussianRandomWalk
    from flowmc_b200.resource.kernel.HMC import HMC
    from flowmc_b200.resource.kernel.MALA import MALA
    from oracle import local as olocal, targets as O
    from test_gpu_local import _run_gpu
    M = np.eye(d, dtype=np.float32)
    k, ok = {"MALA": (MALA(0.3), olocal.make_kernel("MALA", step_size=0.3)),
             "GRW": (GaussianRandomWalk(0.3), olocal.make_kernel("GRW", step_size=0.3)),
             "HMC": (HMC(M, 0.2, 3), olocal.make_kernel("HMC", step_size=0.2, n_leapfrog=3, condition_matrix=M))}[kind]
    key = frandom.PRNGKey(5)
    x0 = frandom.normal(frandom.PRNGKey(6), (n_chains, d))
    new_key, res, last, strat = _run_gpu(k, T.iso_gaussian(0.5, None), None, d, key, x0, n_steps)
    o_key, o_pos, o_lp, o_acc, o_last = olocal.take_serial_steps(key, x0.cpu().numpy(), "iso_gaussian",
                                                                 O.IsoGaussian.pack(d, 0.5), ok, n_steps)
    assert np.array_equal(new_key, o_key)
    assert res["positions"].data.shape == (n_chains, n_steps, d)
    ga = res["acceptance"].data.cpu().numpy()
    if np.array_equal(ga, o_acc):   # (a near-tie could flip a flag; these seeds have none)
        assert_close(res["positions"].data.cpu().numpy(), o_pos, "positions", rtol=3e-4)
        assert_close(res["log_prob"].data.cpu().numpy(), o_lp, "log_prob", rtol=3e-4)
        assert_close(last.cpu().numpy(), o_last, "last", rtol=3e-4)
    else:
        assert (ga != o_acc).mean() < 0.1


def test_global_step_single_proposal(cuda):
    """TakeGroupSteps with n_steps = 1 and a single chain."""
    from flowmc_b200 import random as frandom, targets as T
    from flowmc_b200.resource.buffers import Buffer
    from flowmc_b200.resource.kernel.NF_proposal import NFProposal
    from flowmc_b200.resource.logPDF import LogPDF
    from flowmc_b200.resource.states import State
    from flowmc_b200.strategy.take_steps import TakeGroupSteps
    d = 5
    m = model_from_params(random_params(4, d, 2, [32, 32], 8))
    for n_chains in (1, 3):
        res = {"p": Buffer("p", (n_chains, 2, d), 1), "l": Buffer("l", (n_chains, 2), 1), "a": Buffer("a", (n_chains, 2), 1),
               "s": State({"p": "p", "l": "l", "a": "a"}, "s"), "k": NFProposal(m),
               "logpdf": LogPDF(T.dual_moon(), n_dims=d)}
        strat = TakeGroupSteps("logpdf", "k", "s", ["p", "l", "a"], 1)
        x0 = frandom.normal(frandom.PRNGKey(2), (n_chains, d))
        key, res, last = strat(frandom.PRNGKey(3), res, x0, None)
        torch.cuda.synchronize()
        assert last.shape == (n_chains, d) and strat.current_position == 1
        acc = res["a"].data[:, 0]
        assert set(acc.cpu().numpy().tolist()) <= {0.0, 1.0}
        assert torch.isinf(res["a"].data[:, 1]).all()          # the untouched slot keeps the -inf fill (buffers.py:27)
        stay = acc == 0
        assert torch.equal(last[stay], x0[stay])


@pytest.mark.parametrize("d,hidden", [(130, [64, 64]), (6, [40, 24])])
def test_shapes_outside_the_tensor_core_path_fall_back_to_cuda_cores(cuda, monkeypatch, d, hidden):
    """n_features > 128 or hidden widths that are not multiples of 16: forward, sampling and a training step run on the
    fp32 CUDA-core kernels (chosen per model) even though the tensor-core path is requested."""
    from oracle import flow as oflow, nf
    monkeypatch.setenv("FLOWMC_FLOW_TC", "3")
    p = random_params(13, d, 2, hidden, 8, gain=1.5)
    m = model_from_params(p)
    r = np.random.default_rng(d)
    x = (1.5 * r.standard_normal((70, d))).astype(np.float32)
    lp = m.log_prob(torch.from_numpy(x).cuda())
    o32 = oflow.log_prob(p, x)
    with oflow.precision(np.float64):
        o64 = oflow.log_prob(p, x)
    assert_close(lp.cpu().numpy(), o32, "log_prob", floor=o32 - o64, floor_factor=3.0)
    loss, grad = m.loss_and_grad(torch.from_numpy(x).cuda())
    o_loss, _ = nf.loss_and_grads(p, x)
    assert abs(float(loss) - o_loss) <= 1e-5 * max(1.0, abs(o_loss))
    assert bool(torch.isfinite(grad).all()) and float(grad.abs().max()) > 0
    from oracle import rng
    assert m.sample(rng.PRNGKey(2), 33).shape == (33, d)
