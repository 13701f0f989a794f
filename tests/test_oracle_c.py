"""Pins the C restatement (oracle/c/flowmc_ref.c -- the CPU baseline bench.py times and the checker of the
BASELINE-scale GPU parity tests) to the numpy oracle (oracle/local.py), which restates
src/flowMC/strategy/take_steps.py:60-180 and resource/kernel/{MALA,HMC,Gaussian_random_walk}.py step by step.

Keys and accept flags must be bit-equal; positions / log-probs agree to float32 rounding (the two evaluate the same
fp32 formulas; libm vs numpy transcendental functions and the summation order of numpy reductions differ in the last
bits, and the differences feed back through the chain)."""
import numpy as np
import pytest

from oracle import cref, local as olocal, rng, targets as O

TARGETS = ["iso_gaussian", "dual_moon", "ar1_gaussian", "dense_gaussian", "rosenbrock", "gaussian_mixture"]
KINDS = ["MALA", "HMC", "GRW"]


def _pack(name, d):
    rs = np.random.RandomState(5)
    if name == "iso_gaussian":
        return O.IsoGaussian.pack(d, 0.5)
    if name == "dual_moon":
        return O.DualMoon.pack(d, np.arange(d) * 0.1)
    if name == "ar1_gaussian":
        return O.AR1Gaussian.pack(d, 0.9)
    if name == "dense_gaussian":
        P = rs.randn(d, d)
        return O.DenseGaussian.pack(d, P @ P.T / d + np.eye(d))
    if name == "rosenbrock":
        return O.Rosenbrock.pack(d)
    if name == "gaussian_mixture":
        return O.GaussianMixture.pack(d, rs.randn(8, d).astype(np.float32) * 2, 1.0)
    raise KeyError(name)


def _kernel(kind, d, target):
    step = {"MALA": 0.1, "GRW": 0.2, "HMC": 0.05}[kind]
    if target == "rosenbrock":
        step *= 0.2
    kw = dict(step_size=step)
    ckw = dict(step_size=step)
    if kind == "HMC":
        rs = np.random.RandomState(9)
        A = rs.randn(d, d) * 0.2
        M = (A @ A.T + np.eye(d)).astype(np.float32)          # dense inverse mass: exercises the Cholesky matvec
        kw.update(n_leapfrog=4, condition_matrix=M)
        L, colsum = olocal.hmc_setup(M, d)
        ckw.update(n_leapfrog=4, chol=L, colsum=colsum)
    return olocal.make_kernel(kind, **kw), ckw


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("target", TARGETS)
def test_c_port_equals_numpy_oracle(target, kind):
    d, n, T, thin = 12, 41, 45, 1
    data = _pack(target, d)
    key = rng.PRNGKey(42)
    key, sub = rng.split(key)
    x0 = rng.normal(sub, (n, d))
    ok, ckw = _kernel(kind, d, target)
    o_key, o_pos, o_lp, o_acc, o_last, dbg = olocal.take_serial_steps(key, x0, target, data, ok, T, thinning=thin,
                                                                      return_debug=True)
    c_key, c_pos, c_lp, c_acc, c_last, c_ratio, c_logu = cref.take_serial_steps(
        key, x0, target, data, kind, T, thinning=thin, debug=True, **ckw)
    assert np.array_equal(c_key, o_key)
    # log(uniform) depends on the RNG words only: bit-equal words -> values equal to libm-vs-numpy log rounding
    o_logu = np.stack([dbg[t]["log_u"] for t in range(0, T, thin)], axis=1)
    o_ratio = np.stack([dbg[t]["ratio"] for t in range(0, T, thin)], axis=1)
    np.testing.assert_allclose(c_logu, o_logu, rtol=1e-6, atol=1e-7)
    differs = c_acc != o_acc
    if differs.any():   # only at a near-tie of the accept test, and only for a chain's first difference
        first = np.where(differs.any(1), differs.argmax(1), -1)
        for c in np.nonzero(first >= 0)[0]:
            t = first[c]
            assert abs(o_ratio[c, t] - o_logu[c, t]) <= 1e-4 * max(1.0, abs(o_ratio[c, t])), (c, t)
        assert (first >= 0).sum() <= 1
    same = ~differs.any(1)
    assert same.sum() >= n - 1
    scale = max(1.0, float(np.abs(o_pos).max()))
    assert np.abs(c_pos[same] - o_pos[same]).max() <= 3e-5 * scale
    fin = np.isfinite(o_lp[same])
    assert np.array_equal(np.isfinite(c_lp[same]), fin)
    assert np.abs(c_lp[same][fin] - o_lp[same][fin]).max() <= 3e-5 * max(1.0, float(np.abs(o_lp[same][fin]).max()))
    assert np.abs(c_last[same] - o_last[same]).max() <= 3e-5 * scale
    np.testing.assert_allclose(c_ratio[same][:, 0], o_ratio[same][:, 0], rtol=1e-4, atol=1e-4)


def test_c_port_thinning_is_a_stride_of_the_full_run():
    d, n, T = 6, 9, 20
    data = _pack("dual_moon", d)
    key = rng.PRNGKey(5)
    x0 = rng.normal(rng.split(key)[1], (n, d))
    full = cref.take_serial_steps(key, x0, "dual_moon", data, "MALA", T, debug=True)
    thin = cref.take_serial_steps(key, x0, "dual_moon", data, "MALA", T, thinning=3, debug=True)
    assert np.array_equal(full[0], thin[0])
    for a, b in zip((full[1], full[2], full[3], full[5], full[6]), (thin[1], thin[2], thin[3], thin[5], thin[6])):
        assert np.array_equal(a[:, ::3], b)
    assert np.array_equal(thin[4], thin[1][:, -1])          # last = positions[:, -1] of the THINNED block


def test_c_port_rng_words_bit_exact():
    """threefry2x32 KATs (Random123) and the normal() transform against the numpy oracle."""
    assert cref.threefry2x32(0, 0, 0, 0) == (0x6b200159, 0x99ba4efe)
    assert cref.threefry2x32(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff) == (0x1cb996fc, 0xbb002be7)
    assert cref.threefry2x32(0x13198a2e, 0x03707344, 0x243f6a88, 0x85a308d3) == (0xc4923a9c, 0x483df7a0)
    key = rng.PRNGKey(7)
    z_c = cref.normal(key, 1000)
    z_o = rng.normal(key, (1000,))
    np.testing.assert_allclose(z_c, z_o, rtol=2e-6, atol=1e-7)


def test_c_port_chain_offset_is_a_shard_of_the_global_run():
    d, n, T = 8, 24, 10
    data = _pack("ar1_gaussian", d)
    key = rng.PRNGKey(3)
    x0 = rng.normal(rng.split(key)[1], (n, d))
    full = cref.take_serial_steps(key, x0, "ar1_gaussian", data, "MALA", T)
    part = cref.take_serial_steps(key, x0[8:20], "ar1_gaussian", data, "MALA", T, chain_offset=8)
    for a, b in zip(full[1:], part[1:]):
        assert np.array_equal(a[8:20], b)


def test_c_port_thread_count_does_not_change_results():
    d, n, T = 16, 64, 12
    data = _pack("gaussian_mixture", d)
    key = rng.PRNGKey(11)
    x0 = rng.normal(rng.split(key)[1], (n, d))
    n0 = cref.num_threads()
    try:
        cref.set_num_threads(1)
        a = cref.take_serial_steps(key, x0, "gaussian_mixture", data, "MALA", T)
        cref.set_num_threads(max(2, n0))
        b = cref.take_serial_steps(key, x0, "gaussian_mixture", data, "MALA", T)
    finally:
        cref.set_num_threads(n0)
    for u, v in zip(a, b):
        assert np.array_equal(u, v)
