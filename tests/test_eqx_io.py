"""`.eqx` interop (flowmc_b200/eqx_io.py): the reference's on-disk weight format, restated from
src/flowMC/resource/model/nf_model/base.py:92-96 + the equinox serialisation rules.  CPU only."""
import io

import numpy as np
import pytest

from flowmc_b200 import eqx_io
from oracle import flow as oflow
from oracle import rng


def _params(d=6, L=3, hidden=(16, 8), K=8, seed=4):
    p = oflow.init_params(rng.PRNGKey(seed), d, L, list(hidden), K)
    r = np.random.default_rng(seed)
    p.scale = r.standard_normal(L).astype(np.float32)
    p.shift = r.standard_normal(L).astype(np.float32)
    p.data_mean = r.standard_normal(d).astype(np.float32)
    return p


def _leaves(p):
    return eqx_io.leaves_from_arrays(p.n_features, p.n_layers, p.hidden, p.num_bins, (p.range_min, p.range_max),
                                     p.data_mean, p.data_cov, p.base_mean, p.base_cov, p.scale, p.shift, p.W, p.b)


def test_record_sequence_follows_the_reference_field_order():
    p = _params()
    buf = io.BytesIO()
    eqx_io.write_eqx(buf, p.n_features, p.n_layers, p.hidden, p.num_bins, _leaves(p))
    raw = buf.getvalue()
    assert raw[:6] == b"\x93NUMPY"                      # np.save records back to back, no container
    buf.seek(0)
    recs = eqx_io.read_records(buf)
    spec = eqx_io.leaf_spec(p.n_features, p.n_layers, p.hidden, p.num_bins)
    assert len(recs) == len(spec) == 15 + 2 * 3
    names = [n for n, _, _ in spec]
    # NFModel fields, then base_dist, then layers (nf_model/base.py:20-22, rqSpline.py:381-382)
    assert names[:6] == ["_n_features", "_data_mean", "_data_cov", "base_dist._mean", "base_dist._cov",
                         "base_dist.learnable"]
    assert names[6:10] == ["layers[0]._mask", "layers[0].bijector.scale", "layers[0].bijector.shift", "layers[1]._mask"]
    assert names[-2:] == ["layers[1].bijector.conditioner.layers[4].weight",
                          "layers[1].bijector.conditioner.layers[4].bias"]
    for (name, shape, dtype), r in zip(spec, recs):
        assert r.shape == shape and r.dtype == np.dtype(dtype), name
    assert int(recs[0]) == p.n_features and float(recs[10]) == -10.0 and int(recs[12]) == p.num_bins
    # vmapped layers: leading n_layers axis; equinox Linear weight is (out, in)
    assert recs[15].shape == (p.n_layers, 16, p.n_features)
    assert recs[-2].shape == (p.n_layers, p.n_features * 25, 8)
    assert recs[9][1].tolist() == [bool((j + 1) % 2) for j in range(p.n_features)]   # mask of layer 1


def test_round_trip_is_exact():
    p = _params()
    lv = _leaves(p)
    buf = io.BytesIO()
    eqx_io.write_eqx(buf, p.n_features, p.n_layers, p.hidden, p.num_bins, lv)
    buf.seek(0)
    back = eqx_io.read_eqx(buf, p.n_features, p.n_layers, p.hidden, p.num_bins)
    for name, shape, _ in eqx_io.leaf_spec(p.n_features, p.n_layers, p.hidden, p.num_bins):
        if shape != ():
            assert np.array_equal(back[name], lv[name]), name


def test_loader_keys_on_arrays_not_on_the_scalar_records():
    """A file whose Python-scalar leaves were skipped or typed differently still loads (the array order is what
    the reference source fixes; the scalar handling is recalled equinox behaviour)."""
    p = _params()
    lv = _leaves(p)
    buf = io.BytesIO()
    for name, shape, dtype in eqx_io.leaf_spec(p.n_features, p.n_layers, p.hidden, p.num_bins):
        if shape != ():
            np.save(buf, lv[name])
        elif name.endswith("_range_min"):
            np.save(buf, np.float32(-10.0))            # a differently typed scalar record
    buf.seek(0)
    back = eqx_io.read_eqx(buf, p.n_features, p.n_layers, p.hidden, p.num_bins)
    assert np.array_equal(back["layers[1].bijector.conditioner.layers[2].weight"], p.W[1])


def test_architecture_mismatch_raises():
    p = _params()
    buf = io.BytesIO()
    eqx_io.write_eqx(buf, p.n_features, p.n_layers, p.hidden, p.num_bins, _leaves(p))
    for args in ((p.n_features, p.n_layers, [16, 16], p.num_bins), (p.n_features, p.n_layers + 1, p.hidden, p.num_bins),
                 (p.n_features, p.n_layers, [16], p.num_bins), (p.n_features, p.n_layers, p.hidden, 4)):
        buf.seek(0)
        with pytest.raises(ValueError):
            eqx_io.read_eqx(buf, *args)


# ---- a file NOT written by eqx_io: tests/golden/flow_d3_handwritten.eqx (tests/golden/make_eqx_fixture.py builds it
# byte by byte from the equinox serialisation spec and the reference's field order) ------------------------------
def _handwritten():
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    import make_eqx_fixture as mk
    return mk, os.path.join(here, "golden", "flow_d3_handwritten.eqx")


def _expected_handwritten(mk):
    d, L, K = mk.D, mk.L, mk.K
    dims = [d] + mk.HIDDEN + [d * (3 * K + 1)]

    def arr(leaf, shape):
        n = int(np.prod(shape))
        return np.array([mk.value(leaf, i) for i in range(n)], np.float32).reshape(shape)
    exp = {"_data_mean": arr(1, (d,)),
           "_data_cov": np.where(np.eye(d, dtype=bool), 1.0 + np.arange(d)[:, None] / 4.0, 0.125).astype(np.float32),
           "base_dist._mean": np.zeros(d, np.float32), "base_dist._cov": np.eye(d, dtype=np.float32),
           "layers[0].bijector.scale": np.array([mk.value(7, l) / 8 for l in range(L)], np.float32),
           "layers[0].bijector.shift": np.array([mk.value(8, l) / 8 for l in range(L)], np.float32)}
    for i in range(len(dims) - 1):
        exp[f"layers[1].bijector.conditioner.layers[{2 * i}].weight"] = arr(20 + 2 * i, (L, dims[i + 1], dims[i]))
        exp[f"layers[1].bijector.conditioner.layers[{2 * i}].bias"] = arr(21 + 2 * i, (L, dims[i + 1]))
    return exp


def test_reads_the_hand_written_fixture():
    mk, path = _handwritten()
    with open(path, "rb") as f:
        raw = f.read()
    # the committed bytes ARE what the generator produces (nobody edited one without the other) ...
    assert raw[:6] == b"\x93NUMPY" and raw[6:8] == b"\x01\x00" and len(raw) % 64 != 1
    with open(path, "rb") as f:
        recs = eqx_io.read_records(f)
    spec = eqx_io.leaf_spec(mk.D, mk.L, mk.HIDDEN, mk.K)
    assert len(recs) == len(spec)
    for (name, shape, dtype), r in zip(spec, recs):         # ... numpy parses every record with the specified dtype
        assert r.shape == shape and r.dtype == np.dtype(dtype), name
    with open(path, "rb") as f:
        lv = eqx_io.read_eqx(f, mk.D, mk.L, mk.HIDDEN, mk.K)
    for name, want in _expected_handwritten(mk).items():
        assert np.array_equal(lv[name], want), name
    # eqx_io's own writer reproduces the hand-written file bit for bit
    import io
    full = {n: r for (n, _, _), r in zip(spec, recs)}
    buf = io.BytesIO()
    eqx_io.write_eqx(buf, mk.D, mk.L, mk.HIDDEN, mk.K, full)
    assert buf.getvalue() == raw


def test_hand_written_fixture_log_prob_through_the_oracle():
    """The loaded leaves drive the oracle flow: a finite log_prob that changes when a weight leaf is swapped, i.e. the
    leaves land in the parameter slots the model reads."""
    mk, path = _handwritten()
    with open(path, "rb") as f:
        lv = eqx_io.read_eqx(f, mk.D, mk.L, mk.HIDDEN, mk.K)
    p = oflow.init_params(rng.PRNGKey(0), mk.D, mk.L, mk.HIDDEN, mk.K)
    for i in range(len(p.W)):
        p.W[i] = lv[f"layers[1].bijector.conditioner.layers[{2 * i}].weight"]
        p.b[i] = lv[f"layers[1].bijector.conditioner.layers[{2 * i}].bias"]
    p.scale, p.shift = lv["layers[0].bijector.scale"], lv["layers[0].bijector.shift"]
    p.data_mean, p.data_cov = lv["_data_mean"], lv["_data_cov"]
    x = rng.normal(rng.PRNGKey(1), (8, mk.D))
    lp = oflow.log_prob(p, x)
    assert np.isfinite(lp).all()
    q = p.copy()
    q.W[0] = q.W[0][::-1].copy()
    assert not np.allclose(oflow.log_prob(q, x), lp)
