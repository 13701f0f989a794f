"""Writes tests/golden/flow_d3_handwritten.eqx BYTE BY BYTE from the equinox serialisation spec, without np.save and
without flowmc_b200/eqx_io.py, so that eqx_io's reader is checked against something other than its own writer.

Spec (equinox 0.11.11 ``eqx.tree_serialise_leaves`` with the default filter, as recalled -- equinox is not
installable here, so this is still NOT a file written by the real library): the leaves of
``jax.tree_util.tree_leaves(model)`` in flattening order, each as one ``.npy`` version-1.0 record, back to back:
    magic b"\\x93NUMPY", version bytes 1, 0, little-endian uint16 header length, then the header -- the literal
    "{'descr': '<f4', 'fortran_order': False, 'shape': (2, 3), }" padded with spaces and a final "\\n" so that
    magic + version + length + header is a multiple of 64 bytes -- then the raw C-order little-endian data.
Python bool / int / float leaves are written like 0-d arrays ('|b1' / '<i8' / '<f8', shape ()); function leaves
(jax.nn.tanh in MLP.layers) are skipped.  A Module flattens to its dataclass fields in declaration order, base
classes first, static fields excluded.  Field order transcribed from the reference source:
    NFModel:                 _n_features, _data_mean, _data_cov            nf_model/base.py:20-22
    MaskedCouplingRQSpline:  base_dist, layers                             nf_model/rqSpline.py:377-378
    Gaussian:                _mean, _cov, learnable                        common.py:257-259
    layers = filter_vmap(make_layer) -> eqx.nn.Sequential([layer1, layer2]) (rqSpline.py:427-443): arrays carry a
                             leading n_layers axis, Python scalars do not
    MaskedCouplingLayer:     _mask, bijector                               common.py:139-140
    ScalarAffine:            scale, shift                                  common.py:212-213
    RQSpline:                _range_min, _range_max, _num_bins, _min_bin_size, _min_knot_slope, conditioner
                                                                           nf_model/rqSpline.py:243-248
    MLP:                     layers = [Linear, tanh, Linear, tanh, Linear] common.py:81,91-107
    eqx.nn.Linear:           weight, bias  (in_features / out_features / use_bias are static)

Values are a closed-form function of (leaf index, element index) -- ``value()`` below -- which the tests recompute.
Run from the repo root:  python tests/golden/make_eqx_fixture.py
"""
import os
import struct

D, L, HIDDEN, K = 3, 2, [4, 4], 4
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "flow_d3_handwritten.eqx")


def value(leaf: int, i: int) -> float:
    """Deterministic float32-exact values: multiples of 1/64 in (-2, 2)."""
    return (((leaf * 37 + i * 11) % 255) - 127) / 64.0


def npy_record(descr: str, shape: tuple, payload: bytes) -> bytes:
    shape_s = "()" if shape == () else "(" + ", ".join(str(s) for s in shape) + ("," if len(shape) == 1 else "") + ")"
    header = "{'descr': '%s', 'fortran_order': False, 'shape': %s, }" % (descr, shape_s)
    pad = 64 - ((10 + len(header) + 1) % 64)
    header = header + " " * (pad % 64) + "\n"
    assert (10 + len(header)) % 64 == 0
    return b"\x93NUMPY" + bytes([1, 0]) + struct.pack("<H", len(header)) + header.encode("latin1") + payload


def f32(leaf, shape):
    n = 1
    for s in shape:
        n *= s
    return npy_record("<f4", shape, struct.pack("<%df" % n, *[value(leaf, i) for i in range(n)]))


def main():
    dims = [D] + HIDDEN + [D * (3 * K + 1)]
    rec = []
    rec.append(npy_record("<i8", (), struct.pack("<q", D)))                                   # _n_features
    rec.append(f32(1, (D,)))                                                                  # _data_mean
    # _data_cov: symmetric positive definite, diag 1 + j/4, off-diagonal 1/8
    cov = [(1.0 + r / 4.0) if r == c else 0.125 for r in range(D) for c in range(D)]
    rec.append(npy_record("<f4", (D, D), struct.pack("<%df" % (D * D), *cov)))
    rec.append(npy_record("<f4", (D,), struct.pack("<%df" % D, *([0.0] * D))))                # base_dist._mean
    eye = [1.0 if r == c else 0.0 for r in range(D) for c in range(D)]
    rec.append(npy_record("<f4", (D, D), struct.pack("<%df" % (D * D), *eye)))                # base_dist._cov
    rec.append(npy_record("|b1", (), bytes([0])))                                             # base_dist.learnable
    rec.append(npy_record("|b1", (L, D), bytes([0] * (L * D))))                               # layers[0]._mask
    rec.append(npy_record("<f4", (L,), struct.pack("<%df" % L, *[value(7, l) / 8 for l in range(L)])))   # scale
    rec.append(npy_record("<f4", (L,), struct.pack("<%df" % L, *[value(8, l) / 8 for l in range(L)])))   # shift
    mask = [(j + l) % 2 for l in range(L) for j in range(D)]
    rec.append(npy_record("|b1", (L, D), bytes(mask)))                                        # layers[1]._mask
    rec.append(npy_record("<f8", (), struct.pack("<d", -10.0)))                               # _range_min
    rec.append(npy_record("<f8", (), struct.pack("<d", 10.0)))                                # _range_max
    rec.append(npy_record("<i8", (), struct.pack("<q", K)))                                   # _num_bins
    rec.append(npy_record("<f8", (), struct.pack("<d", 1e-4)))                                # _min_bin_size
    rec.append(npy_record("<f8", (), struct.pack("<d", 1e-4)))                                # _min_knot_slope
    for i in range(len(dims) - 1):
        rec.append(f32(20 + 2 * i, (L, dims[i + 1], dims[i])))                                # Linear.weight
        rec.append(f32(21 + 2 * i, (L, dims[i + 1])))                                         # Linear.bias
    with open(OUT, "wb") as f:
        f.write(b"".join(rec))
    print(OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
