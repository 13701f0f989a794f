"""Read / write the reference's ``.eqx`` weight files for ``MaskedCouplingRQSpline``.

The reference saves a flow with ``eqx.tree_serialise_leaves(path + ".eqx", model)``
(src/flowMC/resource/model/nf_model/base.py:92-96).  equinox (0.11.11, as recalled -- it is not installable
here) writes every leaf of ``jax.tree_util.tree_leaves(model)`` in flattening order, back to back, in ``.npy``
format: arrays with ``jnp.save`` / ``np.save``, Python ``bool`` / ``int`` / ``float`` leaves with ``np.save`` of the
scalar (0-d ``|b1`` / ``<i8`` / ``<f8``), anything else (the ``jax.nn.tanh`` function leaves of the MLP) is skipped.
An equinox ``Module`` flattens to its dataclass fields in declaration order, base classes first, static fields
(``eqx.nn.Linear.in_features / out_features / use_bias``) excluded.  For the model built at
``rqSpline.py:392-443`` -- the per-layer modules are created under ``eqx.filter_vmap`` (``:442-443``), so every
array leaf carries a leading ``n_layers`` axis -- that gives the record sequence of ``leaf_spec`` below.

The loader does not depend on the scalar records being exactly as recalled: it walks the records in order and
matches the float32 / bool ARRAYS by shape, skipping 0-d records, so a file from the real library loads as long as
the relative order of its arrays is the field order of the reference source (which is in /root/reference).

Pure numpy: usable (and tested) without a GPU.  ``MaskedCouplingRQSpline.save_model / load_model`` call it.
"""
from __future__ import annotations

import io
from typing import BinaryIO, Dict, List, Sequence, Tuple

import numpy as np


def leaf_spec(n_features: int, n_layers: int, hidden: Sequence[int], num_bins: int) -> List[Tuple[str, tuple, str]]:
    """(name, shape, dtype) of every record of the file, in order."""
    d, L = int(n_features), int(n_layers)
    dims = [d] + [int(h) for h in hidden] + [d * (3 * int(num_bins) + 1)]
    spec: List[Tuple[str, tuple, str]] = [
        ("_n_features", (), "<i8"),                      # NFModel fields (nf_model/base.py:20-22)
        ("_data_mean", (d,), "<f4"),
        ("_data_cov", (d, d), "<f4"),
        ("base_dist._mean", (d,), "<f4"),                # Gaussian fields (common.py:257-259)
        ("base_dist._cov", (d, d), "<f4"),
        ("base_dist.learnable", (), "|b1"),
        # layers = eqx.nn.Sequential([MaskedCouplingLayer(ScalarAffine), MaskedCouplingLayer(RQSpline)])
        ("layers[0]._mask", (L, d), "|b1"),              # MaskedCouplingLayer fields (common.py:139-140)
        ("layers[0].bijector.scale", (L,), "<f4"),       # ScalarAffine fields (common.py:212-213)
        ("layers[0].bijector.shift", (L,), "<f4"),
        ("layers[1]._mask", (L, d), "|b1"),
        ("layers[1].bijector._range_min", (), "<f8"),    # RQSpline fields (rqSpline.py:243-248)
        ("layers[1].bijector._range_max", (), "<f8"),
        ("layers[1].bijector._num_bins", (), "<i8"),
        ("layers[1].bijector._min_bin_size", (), "<f8"),
        ("layers[1].bijector._min_knot_slope", (), "<f8"),
    ]
    for i in range(len(dims) - 1):                       # MLP.layers: Linear, tanh, Linear, tanh, Linear (common.py:91-107)
        spec.append((f"layers[1].bijector.conditioner.layers[{2 * i}].weight", (L, dims[i + 1], dims[i]), "<f4"))
        spec.append((f"layers[1].bijector.conditioner.layers[{2 * i}].bias", (L, dims[i + 1]), "<f4"))
    return spec


def leaves_from_arrays(n_features: int, n_layers: int, hidden: Sequence[int], num_bins: int,
                       spline_range: Tuple[float, float], data_mean, data_cov, base_mean, base_cov, scale, shift,
                       W: Sequence[np.ndarray], b: Sequence[np.ndarray]) -> Dict[str, np.ndarray]:
    """Name -> value for every record.  ``W[i]`` is ``[n_layers, out_i, in_i]``, ``b[i]`` is ``[n_layers, out_i]``."""
    d, L = int(n_features), int(n_layers)
    ar = np.arange(d)
    out = {
        "_n_features": np.asarray(d, dtype=np.int64),
        "_data_mean": np.asarray(data_mean, np.float32).reshape(d),
        "_data_cov": np.asarray(data_cov, np.float32).reshape(d, d),
        "base_dist._mean": np.asarray(base_mean, np.float32).reshape(d),
        "base_dist._cov": np.asarray(base_cov, np.float32).reshape(d, d),
        "base_dist.learnable": np.asarray(False),
        "layers[0]._mask": np.zeros((L, d), dtype=bool),                                   # rqSpline.py:435
        "layers[0].bijector.scale": np.asarray(scale, np.float32).reshape(L),
        "layers[0].bijector.shift": np.asarray(shift, np.float32).reshape(L),
        "layers[1]._mask": ((ar[None, :] + np.arange(L)[:, None]) % 2).astype(bool),        # rqSpline.py:434
        "layers[1].bijector._range_min": np.asarray(float(spline_range[0]), dtype=np.float64),
        "layers[1].bijector._range_max": np.asarray(float(spline_range[1]), dtype=np.float64),
        "layers[1].bijector._num_bins": np.asarray(int(num_bins), dtype=np.int64),
        "layers[1].bijector._min_bin_size": np.asarray(1e-4, dtype=np.float64),             # rqSpline.py:270-271
        "layers[1].bijector._min_knot_slope": np.asarray(1e-4, dtype=np.float64),
    }
    for i in range(len(W)):
        out[f"layers[1].bijector.conditioner.layers[{2 * i}].weight"] = np.asarray(W[i], np.float32)
        out[f"layers[1].bijector.conditioner.layers[{2 * i}].bias"] = np.asarray(b[i], np.float32)
    return out


def write_eqx(f: BinaryIO, n_features: int, n_layers: int, hidden: Sequence[int], num_bins: int,
              leaves: Dict[str, np.ndarray]) -> None:
    """Write the records of ``leaf_spec`` (values from ``leaves``) back to back in .npy format."""
    for name, shape, dtype in leaf_spec(n_features, n_layers, hidden, num_bins):
        a = np.asarray(leaves[name])
        if tuple(a.shape) != tuple(shape):
            raise ValueError(f"{name}: shape {a.shape}, expected {shape}")
        np.save(f, np.asarray(a.astype(np.dtype(dtype), copy=False), order="C"), allow_pickle=False)  # (keeps 0-d 0-d)


def read_records(f: BinaryIO) -> List[np.ndarray]:
    """All back-to-back .npy records of a file."""
    recs = []
    while True:
        head = f.read(1)
        if not head:
            break
        f.seek(-1, io.SEEK_CUR)
        recs.append(np.load(f, allow_pickle=False))
    return recs


def read_eqx(f: BinaryIO, n_features: int, n_layers: int, hidden: Sequence[int], num_bins: int) -> Dict[str, np.ndarray]:
    """Name -> array for the ARRAY leaves of a ``.eqx`` file of a flow with this architecture (the architecture
    comes from the object being loaded into, exactly like ``eqx.tree_deserialise_leaves(path, like)``).  0-d records
    are not required to be present or typed as recalled; array records must appear in field order with the
    expected shapes -- anything else raises ``ValueError`` (equinox raises on a shape mismatch too)."""
    want = [(n, s, t) for n, s, t in leaf_spec(n_features, n_layers, hidden, num_bins) if s != ()]
    have = [r for r in read_records(f) if r.ndim > 0]
    if len(have) != len(want):
        raise ValueError(f".eqx file holds {len(have)} array leaves, this architecture has {len(want)}")
    out = {}
    for (name, shape, dtype), r in zip(want, have):
        if tuple(r.shape) != tuple(shape):
            raise ValueError(f".eqx leaf {name}: shape {tuple(r.shape)} in the file, {tuple(shape)} expected")
        kind = np.dtype(dtype).kind
        if (kind == "b") != (r.dtype.kind == "b"):
            raise ValueError(f".eqx leaf {name}: dtype {r.dtype} in the file, {np.dtype(dtype)} expected")
        out[name] = r.astype(np.dtype(dtype)) if kind == "b" else r.astype(np.float32)
    L, d = int(n_layers), int(n_features)
    expect = ((np.arange(d)[None, :] + np.arange(L)[:, None]) % 2).astype(bool)
    if not np.array_equal(out["layers[1]._mask"], expect) or out["layers[0]._mask"].any():
        raise ValueError(".eqx file: coupling masks differ from ((arange(d) + layer) % 2) -- not a MaskedCouplingRQSpline "
                         "built by rqSpline.py:434-440")
    return out
