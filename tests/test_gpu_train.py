"""Flow training on the GPU vs the oracle: loss + hand-written backward against float64 autograd of
the restated log_prob, fused clip+AdamW against the restated optax chain, jax.random-compatible
permutation / choice, TrainModel's data selection, and a short end-to-end ``train``.

Mirrors test/unit/test_strategies.py:224-249 (TrainModel runs, returns the right types) and
test/integration/test_normalizingFlow.py (train a few epochs, then sample).
"""
import numpy as np
import pytest
import torch

from flowutil import model_from_params, params_from_model, random_params
from parity import assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["fp32-cuda-cores", "tcgen05-3xtf32"])
def flow_path(request, monkeypatch):
    """Both execution paths of the training forward pass (the tensor-core one also hands its activations to the
    backward kernel instead of having them recomputed)."""
    monkeypatch.setenv("FLOWMC_FLOW_TC", "3" if request.param.startswith("tcgen05") else "0")
    return request.param

GRAD_CASES = [
    # d, layers, hidden, bins, n
    (5, 4, [32, 32], 8, 100),
    (32, 2, [128, 128], 8, 150),
    (64, 2, [128, 128], 8, 70),
    (3, 2, [17, 9], 4, 33),
    (2, 2, [16], 16, 64),
    (7, 2, [8, 8, 8], 8, 200),
    (32, 10, [128, 128], 8, 260),     # the C4 flow (BASELINE.json configs[3]): depth 10
    (64, 8, [128, 128], 8, 200),      # the C5 flow (configs[4]): depth 8, d = 64
]


def _device_flat_to_oracle(m, flat):
    """Re-pack a device-layout flat vector (with alignment padding) into the oracle's dict layout."""
    mm = m.clone()
    mm.params.copy_(flat)
    q = params_from_model(mm)
    return dict(W=q.W, b=q.b, scale=q.scale, shift=q.shift)


@pytest.mark.parametrize("d,L,hidden,K,n", GRAD_CASES)
def test_loss_and_grad_match_float64_autograd(cuda, d, L, hidden, K, n):
    from oracle import nf
    p = random_params(17, d, L, hidden, K, gain=2.0, affine=0.1)
    p.base_cov = (p.base_cov * np.float32(0.98)).astype(np.float32)
    m = model_from_params(p)
    r = np.random.default_rng(1)
    x = (2.0 * r.standard_normal((n, d))).astype(np.float32)
    x[0, 0] = 40.0        # whitened value beyond the spline range: linear tail branch of the backward
    loss, grad = m.loss_and_grad(torch.from_numpy(x).cuda())
    o_loss, og = nf.loss_and_grads(p, x)
    assert abs(float(loss.item()) - o_loss) <= 1e-5 * max(1.0, abs(o_loss))
    g = _device_flat_to_oracle(m, grad)
    for name in ("W", "b"):
        for i in range(len(og[name])):
            ref = og[name][i]
            tol = 2e-4 * np.abs(ref).max() + 1e-7
            err = np.abs(g[name][i] - ref).max()
            assert err <= tol, f"d{name}[{i}]: max err {err:.3e} > {tol:.3e}"
    for name in ("scale", "shift"):
        assert_close(g[name], og[name], f"d{name}", rtol=2e-4)
    # non-trainable tail and padding get exactly zero gradient
    mm = m.clone()
    mm.params.copy_(grad)
    assert float(mm.data_mean.abs().max()) == 0.0 and float(mm.data_cov.abs().max()) == 0.0
    assert float(mm.base_cov.abs().max()) == 0.0


def test_loss_grad_row_index_and_batch_slices(cuda):
    """idx gathers rows; gradients of two half batches scaled by 1/n_total add up to the full batch's
    (the data-parallel identity the all-reduce relies on)."""
    p = random_params(3, 5, 3, [32, 32], 8, gain=2.0)
    m = model_from_params(p)
    r = np.random.default_rng(2)
    x = torch.from_numpy((2.0 * r.standard_normal((300, 5))).astype(np.float32)).cuda()
    idx = torch.from_numpy(r.permutation(300)[:128].astype(np.int32)).cuda()
    l_full, g_full = [t.clone() for t in m.loss_and_grad(x, idx)]
    l_ref, g_ref = [t.clone() for t in m.loss_and_grad(x[idx.long()].contiguous())]
    assert torch.allclose(l_full, l_ref, rtol=1e-6)
    assert torch.allclose(g_full, g_ref, rtol=1e-4, atol=1e-7)
    la, ga = [t.clone() for t in m.loss_and_grad(x, idx[:64], n_global=128)]
    lb, gb = [t.clone() for t in m.loss_and_grad(x, idx[64:], n_global=128)]
    assert torch.allclose(la + lb, l_full, rtol=1e-5)
    assert torch.allclose(ga + gb, g_full, rtol=1e-4, atol=1e-7)


def test_clip_adamw_matches_optax_restatement(cuda):
    from flowmc_b200.resource.optimizer import ClipAdamW, OptState
    from flowmc_b200.resource.model.nf_model.base import _TrainScratch
    from oracle import nf
    p = random_params(5, 5, 2, [16, 16], 8)
    m = model_from_params(p)
    n = m.params.numel()
    r = np.random.default_rng(0)
    opt = ClipAdamW(learning_rate=5e-3)
    st = OptState(n, m.params.device)
    ost = nf.AdamWState(n)
    flat = m.params.cpu().numpy().copy()
    sc = _TrainScratch(m, 0, 1)
    for step, gscale in enumerate((5.0, 0.01, 1.0)):       # clipped, unclipped, clipped
        g = (gscale * r.standard_normal(n) / np.sqrt(n)).astype(np.float32)
        sc.grad.copy_(torch.from_numpy(g))
        m._apply_update(opt, st, sc)
        flat, gnorm = nf.clip_adamw(flat, g, ost, 5e-3)
        assert_close(m.params.cpu().numpy(), flat, f"params after step {step}", rtol=2e-6)
        assert_close(st.mu.cpu().numpy(), ost.mu, "mu", rtol=2e-6, scale=float(np.abs(ost.mu).max()))
        assert_close(st.nu.cpu().numpy(), ost.nu, "nu", rtol=2e-6, scale=float(np.abs(ost.nu).max()))
    assert st.count == 3


@pytest.mark.parametrize("n", [1, 7, 1000, 1625, 1626, 50000])
def test_permutation_is_bit_exact(cuda, n):
    import ctypes as C
    from flowmc_b200._lib import check, lib
    from oracle import rng
    key = rng.PRNGKey(9)
    out = torch.empty(n, dtype=torch.int32, device="cuda")
    ws = torch.empty(int(lib.flowmc_random_permutation_workspace_bytes(n)), dtype=torch.uint8, device="cuda")
    check(lib.flowmc_random_permutation(key.ctypes.data_as(C.POINTER(C.c_uint32)), n, out.data_ptr(), ws.data_ptr(),
                                        ws.numel(), torch.cuda.current_stream().cuda_stream))
    assert np.array_equal(out.cpu().numpy(), rng.permutation(key, n))


@pytest.mark.parametrize("pop,m", [(7, 100), (2000, 5000), (65536 * 100, 4096), (3, 1)])
def test_choice_is_bit_exact(cuda, pop, m):
    import ctypes as C
    from flowmc_b200._lib import check, lib
    from oracle import rng
    key = rng.PRNGKey(77)
    out = torch.empty(m, dtype=torch.int32, device="cuda")
    check(lib.flowmc_random_choice(key.ctypes.data_as(C.POINTER(C.c_uint32)), pop, m, out.data_ptr(),
                                   torch.cuda.current_stream().cuda_stream))
    assert np.array_equal(out.cpu().numpy(), rng.choice_with_replacement(key, pop, m))


@pytest.mark.parametrize("n,d", [(1000, 5), (4097, 64), (300, 3), (50, 130), (700, 400), (600, 512)])
def test_mean_cov(cuda, n, d):
    import ctypes as C
    from flowmc_b200._lib import check, lib
    r = np.random.default_rng(4)
    x = (r.standard_normal((n, d)) @ r.standard_normal((d, d)) * 0.5 + 3.0).astype(np.float32)
    xd = torch.from_numpy(x).cuda()
    mean = torch.empty(d, device="cuda")
    cov = torch.empty((d, d), device="cuda")
    sc = torch.empty(max(d, 256), device="cuda")
    check(lib.flowmc_data_mean_cov(xd.data_ptr(), n, d, mean.data_ptr(), cov.data_ptr(), sc.data_ptr(),
                                   torch.cuda.current_stream().cuda_stream))
    np.testing.assert_allclose(mean.cpu().numpy(), x.mean(0), rtol=2e-5, atol=1e-5)
    ref = np.cov(x.T.astype(np.float64))
    np.testing.assert_allclose(cov.cpu().numpy(), ref, rtol=1e-4, atol=1e-4 * np.abs(ref).max())


def test_select_training_data_matches_oracle(cuda):
    from flowmc_b200.strategy.train_model import TrainModel
    from oracle import nf, rng
    n_chains, n_total, d, filled = 12, 30, 4, 17
    r = np.random.default_rng(8)
    buf = np.full((n_chains, n_total, d), -np.inf, np.float32)
    buf[:, :filled] = r.standard_normal((n_chains, filled, d)).astype(np.float32)
    tm = TrainModel("model", "positions", "optimizer", n_max_examples=500, history_window=10)
    key = rng.PRNGKey(3)
    k2, data = tm.select_training_data(key, torch.from_numpy(buf).cuda())
    o_key, o_train_key, o_data, o_idx = nf.select_training_data(key, buf, 500, 10)
    assert np.array_equal(data.cpu().numpy(), o_data)
    assert np.array_equal(rng.split(k2, 2)[0], o_key) and np.array_equal(rng.split(k2, 2)[1], o_train_key)
    # unequal numbers of finite rows: the reference's reshape cannot work -> explicit error
    buf[3, filled - 1, 0] = np.nan
    with pytest.raises(ValueError):
        tm.select_training_data(key, torch.from_numpy(buf).cuda())


def test_train_matches_oracle_for_a_few_steps(cuda):
    """NFModel.train: data statistics, epoch key schedule, permutation batches, best-model tracking."""
    from flowmc_b200.resource.optimizer import Optimizer
    from oracle import flow as oflow, nf, rng
    d, L, hidden, K = 3, 2, [16, 16], 8
    key = rng.PRNGKey(2)
    p = oflow.init_params(key, d, L, hidden, K)
    m = model_from_params(p)
    r = np.random.default_rng(5)
    data = (r.standard_normal((700, d)) * np.array([1.0, 2.0, 0.5]) + np.array([0.5, -1.0, 2.0])).astype(np.float32)
    opt = Optimizer(m, learning_rate=5e-3)
    tkey = rng.PRNGKey(4)
    out_key, best, best_state, losses = m.train(tkey, torch.from_numpy(data).cuda(), opt.optim, opt.optim_state,
                                                num_epochs=3, batch_size=256, verbose=False)
    o_key, o_best, o_state, o_losses = nf.train(p, tkey, data, nf.AdamWState(nf.flatten(p).size), 5e-3, 3, 256)
    assert np.array_equal(out_key, o_key)
    assert_close(losses.cpu().numpy(), o_losses, "epoch losses", rtol=2e-4)
    q = params_from_model(best)
    for i in range(len(p.W)):
        assert_close(q.W[i], o_best.W[i], f"W[{i}] after training", rtol=5e-4)
    assert_close(q.data_mean, o_best.data_mean, "data_mean (weight-decayed)", rtol=1e-4)
    assert_close(q.data_cov, o_best.data_cov, "data_cov", rtol=1e-3)
    assert best_state.count == o_state.count == 6
    # functional contract: the input model and optimiser state are untouched
    assert torch.equal(m.params, model_from_params(p).params) and opt.optim_state.count == 0
    assert float(losses[-1]) < float(losses[0])


def test_train_model_strategy_like_reference(cuda):
    """test/unit/test_strategies.py:224-249: TrainModel for 10 epochs on a filled buffer; type checks."""
    from flowmc_b200 import random as frandom
    from flowmc_b200.resource.buffers import Buffer
    from flowmc_b200.resource.model.nf_model.rqSpline import MaskedCouplingRQSpline
    from flowmc_b200.resource.optimizer import Optimizer
    from flowmc_b200.strategy.train_model import TrainModel
    n_chains, n_steps, d = 5, 25, 3
    key = frandom.PRNGKey(42)
    key, sub = frandom.split(key)
    model = MaskedCouplingRQSpline(d, 2, [16, 16], 8, sub)
    buf = Buffer("test_position", (n_chains, n_steps, d), 1)
    key, sub = frandom.split(key)
    buf.update_buffer(frandom.normal(sub, (n_chains, n_steps, d)))
    res = {"test_position": buf, "model": model, "optimizer": Optimizer(model),
           "loss": Buffer("loss", (10,), 0)}
    strat = TrainModel("model", "test_position", "optimizer", loss_buffer_name="loss", n_epochs=10, batch_size=64,
                       n_max_examples=10000, verbose=False)
    assert repr(strat) == "Train model"
    key2, res, pos = strat(key, res, frandom.normal(sub, (n_chains, d)), {})
    assert isinstance(res["model"], MaskedCouplingRQSpline) and res["model"] is not model
    # the optimiser state returned is the BEST epoch's (nf_model/base.py:196-200,210)
    cnt = res["optimizer"].optim_state.count
    assert isinstance(res["optimizer"], Optimizer) and 0 < cnt <= 10 * (10000 // 64) and cnt % (10000 // 64) == 0
    assert torch.isfinite(res["loss"].data).all() and res["loss"].cursor == 10
    s = res["model"].sample(frandom.PRNGKey(1), 1000)
    assert torch.isfinite(s).all()


@pytest.mark.parametrize("n", [16384, 145 * 128, 149 * 128 + 77, 3 * 148 * 128])
def test_large_batches_tensor_core_backward(cuda, flow_path, monkeypatch, n):
    """Batch sizes that reach every reduction path of the tensor-core backward -- 128 tiles + in-kernel reducer
    CTAs (16384), too few spare SMs for reducers (145 tiles: stand-alone reduce kernel), more tiles than SMs
    (persistent CTAs accumulate several tiles, ragged last tile) -- against the fp32 CUDA-core backward, which the
    cases above pin to float64 autograd; plus bit-reproducibility (fixed-order reduction, no atomics)."""
    if not flow_path.startswith("tcgen05"):
        pytest.skip("compares both paths in one go")
    p = random_params(23, 32, 3, [128, 128], 8, gain=2.0, affine=0.1)
    r = np.random.default_rng(5)
    x = torch.from_numpy((1.5 * r.standard_normal((n, 32))).astype(np.float32)).cuda()
    monkeypatch.setenv("FLOWMC_FLOW_TC", "3")
    m_tc = model_from_params(p)
    assert m_tc.desc.tc_terms == 3 or m_tc.tc_terms == 3
    l1, g1 = [t.clone() for t in m_tc.loss_and_grad(x)]
    l2, g2 = [t.clone() for t in m_tc.loss_and_grad(x)]
    assert torch.equal(g1, g2) and torch.equal(l1, l2), "tensor-core gradients are not bit-reproducible"
    monkeypatch.setenv("FLOWMC_FLOW_TC", "0")
    m_cc = model_from_params(p)
    l0, g0 = [t.clone() for t in m_cc.loss_and_grad(x)]
    assert abs(float(l1) - float(l0)) <= 1e-5 * max(1.0, abs(float(l0)))
    a, b = _device_flat_to_oracle(m_tc, g1), _device_flat_to_oracle(m_cc, g0)
    for name in ("W", "b"):
        for i in range(len(a[name])):
            tol = 2e-4 * np.abs(b[name][i]).max() + 1e-7
            err = np.abs(a[name][i] - b[name][i]).max()
            assert err <= tol, f"d{name}[{i}]: max err {err:.3e} > {tol:.3e}"
    for name in ("scale", "shift"):
        assert_close(a[name], b[name], f"d{name}", rtol=2e-4)
