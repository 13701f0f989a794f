"""The BASELINE.json local configs at full scale, and the time-sliced launch path of the persistent local-step kernel
(csrc/local_steps.cuh: the step range cut into segments that run as successive launches, chain state handed over
through a workspace; opt-in through FlowmcLocalParams.force_n_seg since the r02 sweep showed the one-launch grid is
faster on B200).  C2 8192x1000, C3 32768x200 and C5-local 65536x50 are checked on BOTH launch plans.

 * small shapes, slicing FORCED through FlowmcLocalParams.force_n_seg / slots_override: bit-identity with the
   unsliced launch for every kernel kind and lane layout, with n_steps not a multiple of the segment length or of the
   32-step key chunk, thinning that straddles segment boundaries, and an inactive tail group;
 * the BASELINE shapes themselves against the C port of the oracle (pinned to oracle/local.py by
   tests/test_oracle_c.py), with near-tie accounting (tests/parity.py: compare_chains_vec).

Reference: src/flowMC/strategy/take_steps.py:60-144,156-180."""
import ctypes as C

import numpy as np
import pytest
import torch

from parity import compare_chains_vec
from test_gpu_local import _cond_matrix, _make, _run_gpu

pytestmark = pytest.mark.gpu


def _kernel(kind, d, dense=False):
    from flowmc_b200.resource.kernel.Gaussian_random_walk import GaussianRandomWalk
    from flowmc_b200.resource.kernel.HMC import HMC
    from flowmc_b200.resource.kernel.MALA import MALA
    if kind == "MALA":
        return MALA(0.1)
    if kind == "GRW":
        return GaussianRandomWalk(0.2)
    return HMC(_cond_matrix("HMC", "dense" if dense else "diag", d), 0.05, 3)


def _plan(kernel, target, n, d, n_steps):
    from flowmc_b200._lib import check, lib
    from flowmc_b200.resource.logPDF import LogPDF
    dev = torch.device("cuda")
    p, keep = kernel._local_params(d, dev)
    ws = torch.empty(max(256, int(lib.flowmc_local_steps_workspace_bytes(n, d, p.layout_hint))), dtype=torch.uint8,
                     device=dev)
    p.workspace, p.workspace_bytes = ws.data_ptr(), ws.numel()
    out = (C.c_int * 12)()
    check(lib.flowmc_local_steps_plan(kernel.KIND, LogPDF(target, n_dims=d).target.target_id, n, d, n_steps,
                                      C.byref(p), out))
    names = ["layout", "G", "DPL", "VEC", "n_groups", "slots", "ctas_per_sm", "smem_per_cta", "n_seg", "seg_len",
             "n_rounds", "round_size"]
    return dict(zip(names, list(out)))


def _run(kernel, tgt, data, d, key, x0, T_, thin, n_seg, slots, hint=0):
    kernel.force_n_seg, kernel.slots_override = n_seg, slots
    try:
        new_key, res, last, _ = _run_gpu(kernel, tgt, data, d, key, x0, T_, thinning=thin, layout_hint=hint)
    finally:
        kernel.force_n_seg, kernel.slots_override = 0, 0
    return (np.asarray(new_key).copy(), res["positions"].data.clone(), res["log_prob"].data.clone(),
            res["acceptance"].data.clone(), last.clone())


def _assert_identical(a, b, what):
    assert np.array_equal(a[0], b[0]), f"{what}: key_out differs"
    for u, v, nm in zip(a[1:], b[1:], ("positions", "log_probs", "accept flags", "last position")):
        assert torch.equal(u, v), f"{what}: {nm} differ between the sliced and the one-launch run"


@pytest.mark.parametrize("kind,tname,d", [("MALA", "ar1_gaussian", 128), ("MALA", "gaussian_mixture", 64),
                                          ("HMC", "rosenbrock", 64), ("HMC", "dense_gaussian", 24),
                                          ("GRW", "iso_gaussian", 2), ("GRW", "rosenbrock", 12),
                                          ("MALA", "dual_moon", 5)])
@pytest.mark.parametrize("n_seg,T_,thin", [(2, 70, 1), (3, 100, 3), (5, 37, 2), (4, 129, 7), (7, 50, 1)])
def test_forced_slicing_is_bit_identical(cuda, kind, tname, d, n_seg, T_, thin):
    from flowmc_b200 import random as frandom
    tgt, data, _ = _make(tname, d)
    k = _kernel(kind, d, dense="dense" in tname)
    n = 37                                     # odd: the last chain group of every layout has an inactive tail
    key = frandom.PRNGKey(100 + n_seg)
    x0 = frandom.normal(frandom.split(key)[1], (n, d))
    plan = None
    k.force_n_seg, k.slots_override = n_seg, 5
    try:
        plan = _plan(k, tgt, n, d, T_)
    finally:
        k.force_n_seg, k.slots_override = 0, 0
    assert plan["n_seg"] > 1 and plan["n_rounds"] > plan["n_seg"] - 1 and plan["round_size"] <= 5
    assert plan["seg_len"] * (plan["n_seg"] - 1) < T_ <= plan["seg_len"] * plan["n_seg"]
    one = _run(k, tgt, data, d, key, x0, T_, thin, -1, 0)
    sliced = _run(k, tgt, data, d, key, x0, T_, thin, n_seg, 5)
    _assert_identical(sliced, one, f"{kind}/{tname} S={n_seg} T={T_} thin={thin}")
    # a round as large as the group count (every segment = one launch) is the other extreme of the schedule
    wide = _run(k, tgt, data, d, key, x0, T_, thin, n_seg, 1 << 20)
    _assert_identical(wide, one, f"{kind}/{tname} S={n_seg} wide rounds")


@pytest.mark.parametrize("hint", list(range(1, 12)))
def test_forced_slicing_every_layout(cuda, hint):
    from flowmc_b200 import random as frandom
    from flowmc_b200.resource.kernel.MALA import MALA
    G, DPL, VEC = [(1, 8, 1), (4, 8, 1), (8, 8, 1), (32, 4, 1), (32, 16, 1), (8, 4, 4), (16, 4, 4), (16, 8, 4),
                   (32, 4, 4), (32, 8, 4), (32, 16, 4)][hint - 1]
    d = (G * DPL if hint % 2 == 0 else G * DPL - 4) if VEC == 4 else max(2, G * DPL - 3)
    n, T_ = 19, 45
    tgt, data, _ = _make("ar1_gaussian", d)
    key = frandom.PRNGKey(hint)
    x0 = frandom.normal(frandom.split(key)[1], (n, d))
    k = MALA(0.1)
    one = _run(k, tgt, data, d, key, x0, T_, 2, -1, 0, hint=hint)
    sliced = _run(k, tgt, data, d, key, x0, T_, 2, 4, 3, hint=hint)
    _assert_identical(sliced, one, f"layout {hint}")


def test_sharded_sliced_run_equals_rows_of_the_full_run(cuda):
    """chain_offset with slicing: the shard's key schedule uses GLOBAL chain indices in every segment."""
    from flowmc_b200 import random as frandom
    from flowmc_b200.resource.kernel.MALA import MALA
    from flowmc_b200.resource.logPDF import LogPDF
    from flowmc_b200.strategy.take_steps import TakeSerialSteps
    from test_gpu_local import _setup
    d, n, T_ = 16, 48, 70
    tgt, data, _ = _make("ar1_gaussian", d)
    key = frandom.PRNGKey(3)
    x0 = frandom.normal(frandom.split(key)[1], (n, d))
    k = MALA(0.1)
    full = _run(k, tgt, data, d, key, x0, T_, 1, -1, 0)
    for a, b in ((0, 16), (16, 48)):
        res = _setup(b - a, d, T_)
        ks = MALA(0.1)
        ks.force_n_seg, ks.slots_override = 3, 4
        res["kernel"] = ks
        res["logpdf"] = LogPDF(tgt, n_dims=d)
        s = TakeSerialSteps("logpdf", "kernel", "sampler_state", ["positions", "log_prob", "acceptance"], T_)
        s.set_chain_shard(a, n)
        _, res, last = s(key, res, x0[a:b], data)
        assert torch.equal(res["positions"].data, full[1][a:b])
        assert torch.equal(res["acceptance"].data, full[3][a:b])
        assert torch.equal(last, full[4][a:b])


# ---- the BASELINE.json local configs, exactly, against the C port -------------------------------------------------
def _baseline_case(name):
    from flowmc_b200 import targets as T
    from flowmc_b200.resource.kernel.HMC import HMC
    from flowmc_b200.resource.kernel.MALA import MALA
    from oracle import local as olocal, targets as O
    if name == "C2":      # 128-D AR(1) Gaussian, 8192 chains, MALA 0.1, 1000 steps
        d = 128
        return dict(d=d, n=8192, T=1000, kernel=MALA(0.1), kind="MALA", target=T.ar1_gaussian(0.9), tname="ar1_gaussian",
                    packed=O.AR1Gaussian.pack(d, 0.9), ckw=dict(step_size=0.1))
    if name == "C3":      # 64-D Rosenbrock, 32768 chains, HMC 0.01 x 10 leapfrog, diagonal mass, 200 steps
        d = 64
        M = np.diag(np.linspace(0.5, 2.0, d)).astype(np.float32)
        L, colsum = olocal.hmc_setup(M, d)
        return dict(d=d, n=32768, T=200, kernel=HMC(M, 0.01, 10), kind="HMC", target=T.rosenbrock(), tname="rosenbrock",
                    packed=O.Rosenbrock.pack(d), ckw=dict(step_size=0.01, n_leapfrog=10, chol=L, colsum=colsum))
    if name == "C5-local":  # 64-D 8-component mixture, 65536 chains, MALA 0.1, 50 steps per call
        d = 64
        mu = np.zeros((8, d), np.float32)
        for i in range(8):
            mu[i, i] = 3.0 if i % 2 == 0 else -3.0
        return dict(d=d, n=65536, T=50, kernel=MALA(0.1), kind="MALA", target=T.gaussian_mixture(mu, 1.0),
                    tname="gaussian_mixture", packed=O.GaussianMixture.pack(d, mu, 1.0), ckw=dict(step_size=0.1))
    raise KeyError(name)


@pytest.mark.parametrize("n_seg", [0, 4], ids=["one-launch", "sliced-x4"])
@pytest.mark.parametrize("name", ["C2", "C3", "C5-local"])
def test_baseline_config_against_c_port(cuda, name, n_seg):
    """n_seg = 0: the default plan (what bench.py and the Sampler run); 4: four time segments in resident waves."""
    from flowmc_b200 import random as frandom
    from oracle import cref
    c = _baseline_case(name)
    d, n, T_ = c["d"], c["n"], c["T"]
    c["kernel"].force_n_seg = n_seg
    plan = _plan(c["kernel"], c["target"], n, d, T_)
    assert plan["n_groups"] > plan["slots"], f"{name}: expected more chain groups than resident slots, got {plan}"
    if n_seg > 1:
        assert plan["n_seg"] == n_seg and plan["n_rounds"] > n_seg and plan["round_size"] == plan["slots"], plan
    else:
        assert plan["n_seg"] == 1 and plan["n_rounds"] == 1, plan
    key = frandom.PRNGKey(1)
    x0 = frandom.normal(frandom.split(frandom.PRNGKey(0))[1], (n, d))
    new_key, res, last, _ = _run_gpu(c["kernel"], c["target"], None, d, key, x0, T_)
    gp = res["positions"].data.cpu().numpy()
    gl = res["log_prob"].data.cpu().numpy()
    ga = res["acceptance"].data.cpu().numpy()
    glast = last.cpu().numpy()
    del res
    torch.cuda.empty_cache()
    cref.set_num_threads(0 or __import__("os").cpu_count())
    o_key, o_pos, o_lp, o_acc, o_last, ratio, logu = cref.take_serial_steps(
        key, x0.cpu().numpy(), c["tname"], c["packed"], c["kind"], T_, debug=True, **c["ckw"])
    assert np.array_equal(np.asarray(new_key), o_key)
    assert set(np.unique(ga)) <= {0.0, 1.0}
    nd, worst = compare_chains_vec((gp, gl, ga), (o_pos, o_lp, o_acc), ratio, logu, what=name)
    same = (ga == o_acc).all(axis=1)
    assert same.mean() >= 0.99
    scale = max(1.0, float(np.abs(o_last).max()))
    assert np.abs(glast[same] - o_last[same]).max() <= 3e-4 * scale
    assert np.array_equal(glast, gp[:, -1])
    print(f"{name}: plan {plan}; {nd} of {n} chains diverged at a near-tie; max |dx| before divergence {worst:.2e}; "
          f"acceptance {ga.mean():.3f}")
