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
