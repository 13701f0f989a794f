"""Local kernels (MALA / HMC / GRW) through TakeSerialSteps vs the oracle on the same seeds.

Mirrors the reference's test strategy for this path (test/unit/test_kernels.py,
test/unit/test_strategies.py:117-184) and adds what the reference cannot pin: step-by-step
parity of positions, log-probs and accept flags with the CPU restatement.
"""
import numpy as np
import pytest
import torch

from parity import assert_close, compare_chains, teacher_forced

pytestmark = pytest.mark.gpu


def _setup(n_chains, d, n_total):
    from flowmc_b200.resource.buffers import Buffer
    from flowmc_b200.resource.states import State
    res = {
        "positions": Buffer("positions", (n_chains, n_total, d), 1),
        "log_prob": Buffer("log_prob", (n_chains, n_total), 1),
        "acceptance": Buffer("acceptance", (n_chains, n_total), 1),
        "sampler_state": State({"positions": "positions", "log_prob": "log_prob", "acceptance": "acceptance"},
                               name="sampler_state"),
    }
    return res


def _run_gpu(kernel, target, data, d, key, x0, n_steps, thinning=1, n_total=None, cursor=0, layout_hint=0):
    from flowmc_b200.resource.logPDF import LogPDF
    from flowmc_b200.strategy.take_steps import TakeSerialSteps
    n = x0.shape[0]
    n_out = len(range(0, n_steps, thinning))
    n_total = n_total or (cursor + n_out)
    res = _setup(n, d, n_total)
    res["kernel"] = kernel
    kernel.layout_hint = layout_hint
    res["logpdf"] = LogPDF(target, n_dims=d)
    strat = TakeSerialSteps("logpdf", "kernel", "sampler_state", ["positions", "log_prob", "acceptance"],
                            n_steps, thinning=thinning)
    strat.set_current_position(cursor)
    new_key, res, last = strat(key, res, x0, data)
    torch.cuda.synchronize()
    return new_key, res, last, strat


def _kernels():
    from flowmc_b200.resource.kernel.Gaussian_random_walk import GaussianRandomWalk
    from flowmc_b200.resource.kernel.HMC import HMC
    from flowmc_b200.resource.kernel.MALA import MALA
    return MALA, HMC, GaussianRandomWalk


CONFIGS = [
    # name, kind, kwargs, target factory, oracle target, oracle pack, d, n_chains, n_steps
    ("mala-dualmoon-d5", "MALA", dict(step_size=0.1), "dual_moon", 5, 20, 40),
    ("mala-ar1-d128", "MALA", dict(step_size=0.1), "ar1_gaussian", 128, 37, 70),
    ("mala-mix-d64", "MALA", dict(step_size=0.2), "gaussian_mixture", 64, 50, 40),
    ("grw-iso-d2", "GRW", dict(step_size=0.7), "iso_gaussian", 2, 100, 64),
    ("grw-rosen-d12", "GRW", dict(step_size=0.05), "rosenbrock", 12, 33, 33),
    ("hmc-rosen-d64", "HMC", dict(step_size=0.01, n_leapfrog=10), "rosenbrock", 64, 40, 12),
    ("hmc-iso-d5", "HMC", dict(step_size=0.3, n_leapfrog=5), "iso_gaussian", 5, 64, 40),
    ("hmc-dense-d24", "HMC", dict(step_size=0.1, n_leapfrog=3), "dense_gaussian", 24, 20, 20),
]


def _make(name_t, d):
    from flowmc_b200 import targets as T
    from oracle import targets as O
    rs = np.random.RandomState(5)
    if name_t == "dual_moon":
        return T.dual_moon("data"), {"data": np.arange(d)}, O.DualMoon.pack(d, np.arange(d))
    if name_t == "ar1_gaussian":
        return T.ar1_gaussian(0.9), None, O.AR1Gaussian.pack(d, 0.9)
    if name_t == "gaussian_mixture":
        mu = rs.randn(8, d).astype(np.float32) * 2
        return T.gaussian_mixture(mu, 1.0), None, O.GaussianMixture.pack(d, mu, 1.0)
    if name_t == "iso_gaussian":
        return T.iso_gaussian(0.5, None), None, O.IsoGaussian.pack(d, 0.5)
    if name_t == "rosenbrock":
        return T.rosenbrock(), None, O.Rosenbrock.pack(d)
    if name_t == "dense_gaussian":
        P = rs.randn(d, d)
        P = P @ P.T / d + np.eye(d)
        return T.dense_gaussian(P), None, O.DenseGaussian.pack(d, P)
    raise KeyError(name_t)


def _cond_matrix(kind, name, d):
    if kind != "HMC":
        return None
    if "dense" in name:
        rs = np.random.RandomState(9)
        A = rs.randn(d, d) * 0.2
        return (A @ A.T + np.eye(d)).astype(np.float32)      # dense inverse-mass: exercises the matvec path
    return np.diag(np.linspace(0.5, 2.0, d)).astype(np.float32)


@pytest.mark.parametrize("cfg", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_take_serial_steps_parity(cuda, cfg):
    from flowmc_b200 import random as frandom
    from oracle import local as olocal
    from oracle import rng
    name, kind, kw, tname, d, n, T_ = cfg
    MALA, HMC, GRW = _kernels()
    tgt, data, packed = _make(tname, d)
    key = frandom.PRNGKey(42)
    key, sub = frandom.split(key)
    x0 = frandom.normal(sub, (n, d))
    M = _cond_matrix(kind, name, d)
    if kind == "MALA":
        k = MALA(**kw)
    elif kind == "GRW":
        k = GRW(**kw)
    else:
        k = HMC(condition_matrix=M, **kw)
    new_key, res, last, strat = _run_gpu(k, tgt, data, d, key, x0, T_)
    okw = dict(kw)
    if kind == "HMC":
        okw["condition_matrix"] = M
    ok = olocal.make_kernel(kind, **okw)
    x0_o = rng.normal(sub, (n, d))
    np.testing.assert_allclose(x0.cpu().numpy(), x0_o, rtol=3e-6, atol=1e-7)
    o_key, o_pos, o_lp, o_acc, o_last, dbg = olocal.take_serial_steps(
        key, x0.cpu().numpy(), tname, packed, ok, T_, return_debug=True)
    assert np.array_equal(new_key, o_key)
    gp = res["positions"].data.cpu().numpy()
    gl = res["log_prob"].data.cpu().numpy()
    ga = res["acceptance"].data.cpu().numpy()
    nd = compare_chains((gp, gl, ga), (o_pos, o_lp, o_acc), dbg)
    if nd == 0:
        assert np.array_equal(ga, o_acc)
        assert_close(last.cpu().numpy(), o_last, "last position", rtol=3e-4)
    teacher_forced(k, res["logpdf"], data, dbg, sorted({0, 1, T_ // 3, (2 * T_) // 3, T_ - 1}))
    assert strat.current_position == T_
    assert set(np.unique(ga)) <= {0.0, 1.0}


@pytest.mark.parametrize("hint", list(range(1, 12)))
def test_every_layout_gives_the_same_chains(cuda, hint):
    """All lane layouts of the persistent kernel are the same algorithm: identical flags."""
    from flowmc_b200 import random as frandom
    from flowmc_b200.resource.kernel.MALA import MALA
    from oracle import local as olocal
    G, DPL, VEC = [(1, 8, 1), (4, 8, 1), (8, 8, 1), (32, 4, 1), (32, 16, 1), (8, 4, 4), (16, 4, 4), (16, 8, 4),
                   (32, 4, 4), (32, 8, 4), (32, 16, 4)][hint - 1]
    # float4 layouts: even hints run the exact variant (d == G*DPL), odd ones the padded variant
    d = (G * DPL if hint % 2 == 0 else G * DPL - 4) if VEC == 4 else max(2, G * DPL - 3)
    n, T_ = 19, 37
    tgt, data, packed = _make("ar1_gaussian", d)
    key = frandom.PRNGKey(hint)
    x0 = frandom.normal(frandom.split(key)[1], (n, d))
    new_key, res, last, _ = _run_gpu(MALA(0.1), tgt, data, d, key, x0, T_, layout_hint=hint)
    ok = olocal.make_kernel("MALA", step_size=0.1)
    o_key, o_pos, o_lp, o_acc, o_last, dbg = olocal.take_serial_steps(key, x0.cpu().numpy(), "ar1_gaussian", packed,
                                                                      ok, T_, return_debug=True)
    compare_chains((res["positions"].data.cpu().numpy(), res["log_prob"].data.cpu().numpy(),
                    res["acceptance"].data.cpu().numpy()), (o_pos, o_lp, o_acc), dbg)


def test_thinning_cursor_and_untouched_slots(cuda):
    from flowmc_b200 import random as frandom
    from flowmc_b200.resource.kernel.MALA import MALA
    from oracle import local as olocal
    d, n, T_, thin, cursor, n_total = 8, 13, 50, 3, 4, 40
    tgt, data, packed = _make("iso_gaussian", d)
    key = frandom.PRNGKey(1)
    x0 = frandom.normal(frandom.split(key)[1], (n, d))
    new_key, res, last, strat = _run_gpu(MALA(0.5), tgt, data, d, key, x0, T_, thinning=thin, n_total=n_total,
                                         cursor=cursor)
    ok = olocal.make_kernel("MALA", step_size=0.5)
    _, o_pos, o_lp, o_acc, o_last, dbg = olocal.take_serial_steps(key, x0.cpu().numpy(), "iso_gaussian", packed, ok,
                                                                  T_, thinning=thin, return_debug=True)
    n_out = o_pos.shape[1]
    assert n_out == 17
    gp = res["positions"].data.cpu().numpy()
    ga = res["acceptance"].data.cpu().numpy()
    gl = res["log_prob"].data.cpu().numpy()
    assert np.isneginf(gp[:, :cursor]).all() and np.isneginf(gp[:, cursor + n_out:]).all()
    assert np.isneginf(ga[:, :cursor]).all() and np.isneginf(gl[:, cursor + n_out:]).all()
    assert np.array_equal(ga[:, cursor:cursor + n_out], o_acc)
    assert_close(gp[:, cursor:cursor + n_out], o_pos, "thinned positions", rtol=3e-4)
    assert_close(last.cpu().numpy(), o_last, "last = positions[:, -1] of the thinned block", rtol=3e-4)
    assert torch.equal(last, res["positions"].data[:, cursor + n_out - 1])
    assert strat.current_position == cursor + T_ // thin


def test_chain_sharding_is_bit_identical(cuda):
    """Global-index keys: running chains [a,b) as a shard equals rows [a,b) of the full run."""
    from flowmc_b200 import random as frandom
    from flowmc_b200.resource.kernel.MALA import MALA
    from flowmc_b200.resource.logPDF import LogPDF
    from flowmc_b200.strategy.take_steps import TakeSerialSteps
    d, n, T_ = 16, 48, 25
    tgt, data, _ = _make("ar1_gaussian", d)
    key = frandom.PRNGKey(3)
    x0 = frandom.normal(frandom.split(key)[1], (n, d))
    _, res_full, last_full, _ = _run_gpu(MALA(0.1), tgt, data, d, key, x0, T_)
    for a, b in ((0, 16), (16, 48)):
        res = _setup(b - a, d, T_)
        res["kernel"] = MALA(0.1)
        res["logpdf"] = LogPDF(tgt, n_dims=d)
        s = TakeSerialSteps("logpdf", "kernel", "sampler_state", ["positions", "log_prob", "acceptance"], T_)
        s.set_chain_shard(a, n)
        _, res, last = s(key, res, x0[a:b], data)
        assert torch.equal(res["positions"].data, res_full["positions"].data[a:b])
        assert torch.equal(res["acceptance"].data, res_full["acceptance"].data[a:b])
        assert torch.equal(last, last_full[a:b])


def test_kernel_call_determinism_and_tiny_step_acceptance(cuda):
    """test/unit/test_kernels.py:37-61,111-133,199-239,303-342 re-expressed on the device kernels."""
    from flowmc_b200 import random as frandom
    from flowmc_b200 import targets as T
    from flowmc_b200.resource.logPDF import LogPDF
    MALA, HMC, GRW = _kernels()
    d, n = 2, 100
    logpdf = LogPDF(T.iso_gaussian(0.5, None), n_dims=d)
    key = frandom.PRNGKey(42)
    key, sub = frandom.split(key)
    x0 = frandom.normal(sub, (n, d))
    lp0 = logpdf(x0, None)
    key, sub = frandom.split(key)
    keys = frandom.split(sub, n)
    for k in (MALA(1e-5), GRW(1e-5), HMC(np.eye(d, dtype=np.float32), 1e-7, 5)):
        r1 = k.kernel(keys, x0, lp0, logpdf, None)
        r2 = k.kernel(keys, x0, lp0, logpdf, None)
        for a, b in zip(r1, r2):
            assert torch.equal(a, b)
        assert bool(r1[2].all()), repr(k)
    # single-chain form of the API
    p, l, a = MALA(0.1).kernel(keys[0], x0[0], lp0[0], logpdf, None)
    assert p.shape == (d,) and l.dim() == 0 and a.dim() == 0


def test_kernel_call_matches_oracle_step(cuda):
    from flowmc_b200 import random as frandom
    from flowmc_b200 import targets as T
    from flowmc_b200.resource.logPDF import LogPDF
    from oracle import local as olocal
    from oracle import targets as O
    MALA, HMC, GRW = _kernels()
    d, n = 5, 64
    logpdf = LogPDF(T.dual_moon(), n_dims=d)
    packed = O.DualMoon.pack(d)
    keys = frandom.split(frandom.PRNGKey(8), n)
    x0 = frandom.normal(frandom.PRNGKey(9), (n, d))
    lp0 = logpdf(x0, None)
    M = np.eye(d, dtype=np.float32)
    for k, ok in ((MALA(0.1), olocal.make_kernel("MALA", step_size=0.1)),
                  (GRW(0.1), olocal.make_kernel("GRW", step_size=0.1)),
                  (HMC(M, 0.05, 4), olocal.make_kernel("HMC", step_size=0.05, n_leapfrog=4, condition_matrix=M))):
        p, l, a = k.kernel(keys, x0, lp0, logpdf, None)
        op, ol, oa, info = ok(keys, x0.cpu().numpy(), lp0.cpu().numpy(), "dual_moon", packed)
        near = np.abs(info["ratio"] - info["log_u"]) < 1e-4 * np.maximum(1, np.abs(info["ratio"]))
        same = a.cpu().numpy() == oa
        assert (same | near).all()
        assert_close(p.cpu().numpy()[same], op[same], repr(k))
        assert_close(l.cpu().numpy()[same], ol[same], repr(k) + " lp")


def test_mala_kernel_ignores_the_incoming_log_prob(cuda):
    """MALA.kernel recomputes logpdf(position) and never reads its log_prob argument (MALA.py:59,75,87); HMC and GRW
    do use it (HMC.py:137, Gaussian_random_walk.py:56).  A stale / wrong log_prob must therefore change nothing for
    MALA -- neither the accept decision nor the value returned on rejection -- and must matter for the other two."""
    from flowmc_b200 import random as frandom
    from flowmc_b200 import targets as T
    from flowmc_b200.resource.logPDF import LogPDF
    MALA, HMC, GRW = _kernels()
    d, n = 5, 256
    logpdf = LogPDF(T.dual_moon(), n_dims=d)
    keys = frandom.split(frandom.PRNGKey(8), n)
    x0 = frandom.normal(frandom.PRNGKey(9), (n, d))
    lp0 = logpdf(x0, None)
    wrong = lp0 + 7.5
    k = MALA(0.3)
    p1, l1, a1 = k.kernel(keys, x0, lp0, logpdf, None)
    p2, l2, a2 = k.kernel(keys, x0, wrong, logpdf, None)
    assert torch.equal(p1, p2) and torch.equal(l1, l2) and torch.equal(a1, a2)
    assert 0 < int(a1.sum()) < n                      # both branches of the where() are exercised
    assert torch.equal(l1[~a1], lp0[~a1])             # rejected chains return the FRESH logpdf(position)
    for k in (GRW(0.3), HMC(np.eye(d, dtype=np.float32), 0.05, 4)):
        _, l1, a1 = k.kernel(keys, x0, lp0, logpdf, None)
        _, l2, a2 = k.kernel(keys, x0, wrong, logpdf, None)
        assert torch.equal(l2[~a2], wrong[~a2]), repr(k)   # the incoming value is what a rejection returns
        assert int(a2.sum()) < int(a1.sum()), repr(k)      # ... and it enters the accept ratio


def test_stationarity_unit_gaussian(cuda):
    """test/unit/test_kernels.py:135-184,241-288,344-392: mean ~ 0, var ~ 1 within 3e-2 -- run as
    2048 chains x 2000 steps (the reference runs one chain x 30k-50k steps)."""
    from flowmc_b200 import random as frandom
    from flowmc_b200 import targets as T
    MALA, HMC, GRW = _kernels()
    d, n, T_ = 2, 2048, 2000
    key = frandom.PRNGKey(0)
    x0 = frandom.normal(frandom.split(key)[1], (n, d))
    for k in (MALA(1.0), GRW(1.0), HMC(np.eye(d, dtype=np.float32), 0.5, 5)):
        _, res, _, _ = _run_gpu(k, T.iso_gaussian(0.5, None), None, d, key, x0, T_)
        s = res["positions"].data[:, 200:].reshape(-1, d)
        assert torch.allclose(s.mean(0), torch.zeros(d, device=cuda), atol=3e-2), repr(k)
        assert torch.allclose(s.var(0), torch.ones(d, device=cuda), atol=3e-2), repr(k)
        acc = res["acceptance"].data.mean().item()
        assert 0.2 < acc < 1.0


def test_errors_are_loud(cuda):
    from flowmc_b200 import random as frandom
    from flowmc_b200._lib import FlowmcError
    from flowmc_b200.resource.kernel.MALA import MALA
    tgt, data, _ = _make("iso_gaussian", 4)
    key = frandom.PRNGKey(0)
    x0 = torch.zeros(3, 4, device=cuda)
    with pytest.raises(ValueError):
        _run_gpu(MALA(0.1), tgt, data, 4, key, x0, 10, n_total=5)      # update longer than the buffer
    with pytest.raises(FlowmcError):
        _run_gpu(MALA(0.1), tgt, data, 600, key, torch.zeros(3, 600, device=cuda), 2)   # d > 512
