"""NFProposal global steps (TakeGroupSteps) on the GPU vs the oracle on the same seeds.

Mirrors test/unit/test_strategies.py:251-319 (batched branch n_steps=11 > batch 5, simple branch
n_steps=5) and adds value parity: identical key schedule, accept flags identical except at
near-ties, positions / log-probs within tolerance up to a chain's first near-tie divergence.
"""
import numpy as np
import pytest
import torch

from flowutil import model_from_params, random_params
from parity import assert_close, compare_chains
from test_gpu_local import _make, _setup

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["fp32-cuda-cores", "tcgen05-3xtf32"])
def flow_path(request, monkeypatch):
    """Every test runs on both execution paths of the conditioner GEMMs (shapes outside the tensor-core path's
    limits run the CUDA-core kernels in both)."""
    monkeypatch.setenv("FLOWMC_FLOW_TC", "3" if request.param.startswith("tcgen05") else "0")
    return request.param

# name, target, d, n_chains, n_steps, n_batch_size, thinning, cursor, flow (L, hidden, K), gain
CASES = [
    ("simple-iso-d2", "iso_gaussian", 2, 10, 5, 5, 1, 0, (2, [16, 16], 8), 1.0),
    ("batched-iso-d2", "iso_gaussian", 2, 10, 11, 5, 1, 0, (2, [16, 16], 8), 1.0),
    ("c1-dualmoon-d5", "dual_moon", 5, 20, 10, 100, 1, 3, (4, [32, 32], 8), 1.0),
    ("c5-mix-d64", "gaussian_mixture", 64, 70, 7, 100, 2, 1, (2, [128, 128], 8), 1.5),
    ("batched-thin-d7", "iso_gaussian", 7, 33, 23, 4, 3, 0, (2, [8, 8, 8], 8), 2.0),
]


def _run(case, offset=0, n_shard=None, x0=None):
    from flowmc_b200 import random as frandom
    from flowmc_b200.resource.kernel.NF_proposal import NFProposal
    from flowmc_b200.resource.logPDF import LogPDF
    from flowmc_b200.strategy.take_steps import TakeGroupSteps
    name, tname, d, n, n_steps, bs, thin, cursor, (L, hidden, K), gain = case
    tgt, data, packed = _make(tname, d)
    p = random_params(21, d, L, hidden, K, gain=gain, affine=0.05 * gain, whiten=gain > 1.0)
    m = model_from_params(p)
    key = frandom.PRNGKey(7)
    key, sub = frandom.split(key)
    if x0 is None:
        x0 = frandom.normal(sub, (n, d))
    n_loc = n if n_shard is None else n_shard
    n_out = len(range(0, n_steps, thin))
    res = _setup(n_loc, d, cursor + n_out + 2)
    res["kernel"] = NFProposal(m, n_NFproposal_batch_size=bs)
    res["logpdf"] = LogPDF(tgt, n_dims=d)
    strat = TakeGroupSteps("logpdf", "kernel", "sampler_state", ["positions", "log_prob", "acceptance"], n_steps,
                           thinning=thin)
    strat.set_current_position(cursor)
    if n_shard is not None:
        strat.set_chain_shard(offset, n)
    new_key, res, last = strat(key, res, x0[offset:offset + n_loc], data)
    torch.cuda.synchronize()
    return dict(key=key, new_key=new_key, res=res, last=last, strat=strat, p=p, x0=x0, data=data, packed=packed,
                n_out=n_out)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_take_group_steps_parity(cuda, case):
    from oracle import nf
    name, tname, d, n, n_steps, bs, thin, cursor, flow, gain = case
    g = _run(case)
    o_key, o_pos, o_lp, o_acc, o_last, dbg = nf.take_group_steps(g["key"], g["x0"].cpu().numpy(), g["p"], tname,
                                                                 g["packed"], n_steps, bs, thinning=thin)
    assert np.array_equal(g["new_key"], o_key)
    sl = slice(cursor, cursor + g["n_out"])
    gp = g["res"]["positions"].data[:, sl].cpu().numpy()
    gl = g["res"]["log_prob"].data[:, sl].cpu().numpy()
    ga = g["res"]["acceptance"].data[:, sl].cpu().numpy()
    # untouched slots keep the buffer's -inf initialisation
    assert torch.isinf(g["res"]["log_prob"].data[:, :cursor]).all()
    assert torch.isinf(g["res"]["log_prob"].data[:, cursor + g["n_out"]:]).all()
    steps = dbg["steps"][::thin]
    if thin == 1:
        nd = compare_chains((gp, gl, ga), (o_pos, o_lp, o_acc), steps, max_diverged_frac=0.05)
    else:
        # thinned: a divergence between stored steps shows up later; compare where the flags agree throughout
        same = (ga == o_acc).all(axis=1)
        assert same.mean() > 0.9
        assert_close(gp[same], o_pos[same], "positions", rtol=3e-4)
        assert_close(gl[same], o_lp[same], "log_probs", rtol=3e-4)
        nd = int((~same).sum())
    if nd == 0:
        assert_close(g["last"].cpu().numpy(), o_last, "last position", rtol=3e-4)
    assert g["strat"].current_position == cursor + n_steps // thin
    assert set(np.unique(ga)) <= {0.0, 1.0}
    if gain == 1.0 and tname == "iso_gaussian":
        assert 0.05 < ga.mean() <= 1.0      # near-identity flow on a unit Gaussian: proposals get accepted


def test_proposals_and_log_probs_match_oracle(cuda):
    """NFProposal.kernel called directly (ProposalBase contract) with explicit keys and log_prob."""
    from flowmc_b200.resource.kernel.NF_proposal import NFProposal
    from flowmc_b200.resource.logPDF import LogPDF
    from oracle import nf, rng
    d, n, n_steps = 5, 16, 9
    tgt, data, packed = _make("dual_moon", d)
    p = random_params(3, d, 3, [32, 32], 8, gain=1.0, affine=0.0, whiten=False)
    m = model_from_params(p)
    keys = rng.split(rng.PRNGKey(5), n)
    x = rng.normal(rng.PRNGKey(6), (n, d))
    from oracle.targets import TARGETS
    lp = TARGETS["dual_moon"].logp_grad(x, packed)[0]
    for bs in (100, 4):
        k = NFProposal(m, bs)
        pos, lps, acc = k.kernel(keys, torch.from_numpy(x).cuda(), torch.from_numpy(lp).cuda(),
                                 LogPDF(tgt, n_dims=d), {**data, "n_steps": n_steps})
        o_pos, o_lp, o_acc, dbg = nf.nf_proposal_kernel(p, keys, x, lp, "dual_moon", packed, n_steps, bs)
        assert pos.shape == (n, n_steps, d) and lps.shape == (n, n_steps) and acc.dtype == torch.bool
        compare_chains((pos.cpu().numpy(), lps.cpu().numpy(), acc.cpu().numpy()), (o_pos, o_lp, o_acc), dbg["steps"],
                       max_diverged_frac=0.1)
        # single-chain form
        p1, l1, a1 = k.kernel(keys[3], torch.from_numpy(x[3]).cuda(), torch.tensor(lp[3]).cuda(),
                              LogPDF(tgt, n_dims=d), {**data, "n_steps": n_steps})
        assert torch.equal(p1, pos[3]) and torch.equal(l1, lps[3]) and torch.equal(a1, acc[3])


def test_chain_sharding_is_bit_identical(cuda):
    """Chains [offset, offset+n) of a sharded run equal the same chains of the full run bit for bit."""
    case = CASES[3]
    full = _run(case)
    n = case[3]
    lo = _run(case, offset=0, n_shard=32, x0=full["x0"])
    hi = _run(case, offset=32, n_shard=n - 32, x0=full["x0"])
    for nm in ("positions", "log_prob", "acceptance"):
        a = full["res"][nm].data
        assert torch.equal(a[:32], lo["res"][nm].data) and torch.equal(a[32:], hi["res"][nm].data), nm
    assert np.array_equal(full["new_key"], lo["new_key"]) and np.array_equal(full["new_key"], hi["new_key"])


def test_sample_flow_batched_key_schedule(cuda):
    """sample_flow's batched branch (NF_proposal.py:135-163): n_batch scan iterations, each splitting the key."""
    from flowmc_b200.resource.kernel.NF_proposal import NFProposal
    from oracle import nf, rng
    p = random_params(4, 3, 2, [16, 16], 8, gain=1.0, affine=0.0, whiten=False)
    k = NFProposal(model_from_params(p), 5)
    key = rng.PRNGKey(11)
    x, lp = k.sample_flow(key, 11)
    ox, olp = nf.sample_flow(p, key[None], 11, 5)
    assert_close(x.cpu().numpy(), ox[0], "proposals", rtol=1e-4)
    assert_close(lp.cpu().numpy(), olp[0], "flow log-probs", rtol=1e-4)
