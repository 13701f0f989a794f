"""MaskedCouplingRQSpline forward / inverse / log_prob / sample on the GPU vs the oracle.

Mirrors test/unit/test_nf.py (shapes, forward/inverse consistency) and adds what the reference
cannot pin: values against the CPU restatement within north_star's 1e-5 relative (fp32).
"""
import numpy as np
import pytest
import torch

from flowutil import model_from_params, params_from_model, random_params
from parity import assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["fp32-cuda-cores", "tcgen05-3xtf32"])
def flow_path(request, monkeypatch):
    """Every test runs on both execution paths of the conditioner GEMMs (shapes outside the tensor-core path's
    limits run the CUDA-core kernels in both)."""
    monkeypatch.setenv("FLOWMC_FLOW_TC", "3" if request.param.startswith("tcgen05") else "0")
    return request.param

CASES = [
    # d, layers, hidden, bins, n
    (5, 4, [32, 32], 8, 70),          # C1 quickstart flow
    (32, 3, [128, 128], 8, 200),      # C4 shape (fewer layers)
    (64, 2, [128, 128], 8, 129),      # C5 shape
    (3, 2, [17, 9], 4, 1),            # odd sizes, single row
    (2, 3, [16], 16, 64),             # one hidden layer, 16 bins
    (7, 2, [8, 8, 8], 8, 333),        # three hidden layers
]


def _inputs(seed, n, d, spread=4.0):
    r = np.random.default_rng(seed)
    x = (spread * r.standard_normal((n, d))).astype(np.float32)
    if n > 4:
        x[0, :] = 0.0
        x[1, 0] = 25.0      # linear tails of the spline
        x[2, -1] = -31.0
        x[3, :] = 10.0      # exactly on the range boundary
    return x


# 3xTF32 carries every product to ~2^-22 instead of fp32's 2^-24: where the inverse spline root is
# ill-conditioned the same amplification applies to a ~4x larger input error
def _ff(flow_path):
    return 12.0 if flow_path.startswith("tcgen05") else 3.0


@pytest.mark.parametrize("d,L,hidden,K,n", CASES)
def test_forward_inverse_match_oracle(cuda, flow_path, d, L, hidden, K, n):
    from oracle import flow as oflow
    ff = _ff(flow_path)
    p = random_params(11, d, L, hidden, K)
    m = model_from_params(p)
    x = _inputs(3, n, d)
    y, ld = m.forward(torch.from_numpy(x).cuda())
    oy, old = oflow.forward(p, x)
    with oflow.precision(np.float64):
        oy64, old64 = oflow.forward(p, x)
    assert_close(y.cpu().numpy(), oy, "forward y", floor=oy - oy64, floor_factor=ff)
    assert_close(ld.cpu().numpy(), old, "forward logdet", floor=old - old64, floor_factor=ff)
    xi, ldi = m.inverse(torch.from_numpy(x).cuda())
    ox, oldi = oflow.inverse(p, x)
    with oflow.precision(np.float64):
        ox64, oldi64 = oflow.inverse(p, x)
    assert_close(xi.cpu().numpy(), ox, "inverse x", floor=ox - ox64, floor_factor=ff)
    assert_close(ldi.cpu().numpy(), oldi, "inverse logdet", floor=oldi - oldi64, floor_factor=ff)


@pytest.mark.parametrize("d,L,hidden,K,n", CASES)
def test_log_prob_matches_oracle(cuda, flow_path, d, L, hidden, K, n):
    from oracle import flow as oflow
    p = random_params(5, d, L, hidden, K)
    p.base_cov = (p.base_cov * np.float32(0.97)).astype(np.float32)   # as after AdamW weight decay
    m = model_from_params(p)
    x = _inputs(9, n, d, spread=2.0)
    lp = m.log_prob(torch.from_numpy(x).cuda())
    assert lp.shape == (n,)
    o32 = oflow.log_prob(p, x)
    with oflow.precision(np.float64):
        o64 = oflow.log_prob(p, x)
    assert_close(lp.cpu().numpy(), o32, "log_prob", floor=o32 - o64, floor_factor=_ff(flow_path))


@pytest.mark.parametrize("d,L,hidden,K,n", CASES)
def test_sample_matches_oracle(cuda, flow_path, d, L, hidden, K, n):
    from oracle import flow as oflow, rng
    p = random_params(7, d, L, hidden, K)
    m = model_from_params(p)
    key = rng.PRNGKey(123)
    s = m.sample(key, n)
    assert s.shape == (n, d)
    o32 = oflow.sample(p, key, n)
    with oflow.precision(np.float64):
        o64 = oflow.sample(p, key, n)
    assert_close(s.cpu().numpy(), o32, "sample", floor=o32 - o64, floor_factor=_ff(flow_path))


def test_init_is_bit_exact(cuda):
    """MaskedCouplingRQSpline.__init__ key schedule and draws (rqSpline.py:427-443, common.py:83-107)."""
    from flowmc_b200.resource.model.nf_model.rqSpline import MaskedCouplingRQSpline
    from oracle import flow as oflow, rng
    key = rng.PRNGKey(42)
    m = MaskedCouplingRQSpline(5, 4, [32, 32], 8, key)
    p = oflow.init_params(key, 5, 4, [32, 32], 8)
    q = params_from_model(m)
    for i in range(3):
        assert np.array_equal(q.b[i], p.b[i])          # uniform draws: bit-exact
    assert np.array_equal(q.W[2], p.W[2])
    for i in range(2):                                  # normal draws: same bits, erf_inv polynomial rounds with FMA
        np.testing.assert_allclose(q.W[i], p.W[i], rtol=3e-6, atol=1e-9)
    assert repr(m) == "MaskedCouplingRQSpline with n_features=5, n_layers=4"


def test_forward_inverse_roundtrip_at_init(cuda):
    """test/unit/test_nf.py:38-43 pattern: inverse(forward(x)) == x and logdets cancel (exact inverse
    while ScalarAffine is at its zero initialisation)."""
    from flowmc_b200.resource.model.nf_model.rqSpline import MaskedCouplingRQSpline
    from oracle import rng
    m = MaskedCouplingRQSpline(32, 10, [128, 128], 8, rng.PRNGKey(1))
    x = torch.from_numpy(_inputs(2, 4096, 32, spread=3.0)).cuda()
    y, ld = m.forward(x)
    xb, ldb = m.inverse(y)
    assert torch.allclose(xb, x, rtol=1e-4, atol=1e-4)
    assert torch.allclose(ld, -ldb, rtol=1e-4, atol=1e-3)


def test_shapes_like_reference(cuda):
    """test/unit/test_nf.py:55-73: sample(key, 2) -> (2, 3); log_prob -> (2,)."""
    from flowmc_b200.resource.model.nf_model.rqSpline import MaskedCouplingRQSpline
    from oracle import rng
    m = MaskedCouplingRQSpline(3, 2, [16, 16], 8, rng.PRNGKey(10))
    s = m.sample(rng.PRNGKey(10), 2)
    assert s.shape == (2, 3)
    assert m.log_prob(s).shape == (2,)
    y, ld = m.forward(s[0])
    assert y.shape == (3,) and ld.shape == ()


def test_save_load(cuda, tmp_path):
    from oracle import rng
    from flowmc_b200.resource.model.nf_model.rqSpline import MaskedCouplingRQSpline
    m = MaskedCouplingRQSpline(4, 2, [8, 8], 8, rng.PRNGKey(3))
    m.affine(1).copy_(torch.tensor([0.3, -0.2]))
    m.data_mean.copy_(torch.tensor([1.0, 2.0, 3.0, 4.0]))
    m.save_model(str(tmp_path / "flow"))
    assert (tmp_path / "flow.eqx").exists()            # the reference's file name and format (nf_model/base.py:92-93)
    other = MaskedCouplingRQSpline(4, 2, [8, 8], 8, rng.PRNGKey(99))   # load_model takes the architecture from self
    m2 = other.load_model(str(tmp_path / "flow"))
    assert torch.equal(m.params, m2.params)
    x = torch.randn(50, 4, device="cuda")
    assert torch.equal(m.log_prob(x), m2.log_prob(x))
    with pytest.raises(ValueError):
        MaskedCouplingRQSpline(4, 3, [8, 8], 8, rng.PRNGKey(1)).load_model(str(tmp_path / "flow"))


def test_tensor_core_path_is_really_used(cuda, flow_path):
    """The supported shapes run tcgen05 kernels (plain TF32 shows its ~1e-3 error; 3xTF32 does not)."""
    from oracle import flow as oflow
    if not flow_path.startswith("tcgen05"):
        pytest.skip("CUDA-core parametrisation")
    p = random_params(11, 32, 3, [128, 128], 8)
    x = _inputs(3, 500, 32, spread=2.0)
    ref = oflow.log_prob(p, x)
    m = model_from_params(p)
    assert m.tc_supported()
    lp3 = m.log_prob(torch.from_numpy(x).cuda()).cpu().numpy()
    assert m.desc.tc_terms == 3 and m.desc.tc_image
    m.tc_terms = 1
    lp1 = m.log_prob(torch.from_numpy(x).cuda()).cpu().numpy()
    e3, e1 = np.abs(lp3 - ref).max(), np.abs(lp1 - ref).max()
    assert e3 < 2e-4 * np.abs(ref).max() and e1 > 5 * e3, (e3, e1)
    m.tc_terms = 0
    lp0 = m.log_prob(torch.from_numpy(x).cuda()).cpu().numpy()
    assert m.desc.tc_image is None and np.abs(lp0 - ref).max() < 2e-4 * np.abs(ref).max()
    m2 = model_from_params(random_params(1, 3, 2, [17, 9], 4))
    assert not m2.tc_supported()
