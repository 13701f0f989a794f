"""Regenerates tests/golden/*.npz from the oracle (run from the repo root: python tests/golden/make_golden.py).

The reference itself (flowMC on jax) cannot be imported in this container or on the GPU box (jax,
equinox and optax are absent and there is no network), so these vectors are outputs of the CPU
restatement under oracle/, NOT of flowMC: "parity unpinned" in the task's vocabulary.  What pins the
oracle to the reference is listed in DESIGN.md (Random123 threefry KATs, documented
split(PRNGKey(42)) values, the dual-moon KAT from docs/tutorials/dualmoon.ipynb:65, scipy erfinv,
float64 autograd of the restated log_prob, and the reference's invariant tests).  The fixtures
serve two purposes: they freeze the oracle (tests/test_oracle_golden.py fails if a later edit
changes its results) and give the GPU tests fixed known-answer inputs/outputs that do not
depend on the oracle code being importable.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import flow as oflow, local as olocal, nf, optimization as oopt  # noqa: E402
from oracle import parallel_tempering as opt_pt, realnvp as onvp, rng, targets as otargets  # noqa: E402
from flowutil import random_params  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def flow_case():
    d, L, hidden, K, n = 5, 3, [16, 16], 8, 24
    p = random_params(11, d, L, hidden, K)
    x = (3.0 * np.random.default_rng(3).standard_normal((n, d))).astype(np.float32)
    x[0, 0], x[1, 2] = 17.0, -23.0
    y, ld = oflow.forward(p, x)
    xi, ldi = oflow.inverse(p, x)
    key = rng.PRNGKey(123)
    blob = dict(x=x, fwd_y=y, fwd_logdet=ld, inv_x=xi, inv_logdet=ldi, log_prob=oflow.log_prob(p, x),
                sample_key=key, sample=oflow.sample(p, key, 16), params_flat=nf.flatten(p),
                shape=np.array([d, L, K] + hidden, np.int32))
    loss, g = nf.loss_and_grads(p, x)
    blob.update(loss=np.float32(loss), grad_flat=nf.flatten(g, p))
    np.savez_compressed(os.path.join(OUT, "flow_d5.npz"), **blob)


def init_case():
    key = rng.PRNGKey(42)
    p = oflow.init_params(key, 5, 4, [32, 32], 8)
    np.savez_compressed(os.path.join(OUT, "flow_init_key42.npz"), params_flat=nf.flatten(p),
                        log_prob_zero=oflow.log_prob(p, np.zeros((1, 5), np.float32)))


def local_case():
    d, n, steps = 5, 8, 12
    key = rng.PRNGKey(42)
    ks = rng.split(key, 2)
    x0 = rng.normal(ks[1], (n, d))
    packed = otargets.DualMoon.pack(d, None)
    out = {}
    for kind, kw in (("MALA", dict(step_size=0.1)), ("GRW", dict(step_size=0.3)),
                     ("HMC", dict(step_size=0.05, n_leapfrog=4, condition_matrix=np.eye(d, dtype=np.float32)))):
        k = olocal.make_kernel(kind, **kw)
        nk, pos, lp, acc, last = olocal.take_serial_steps(ks[0], x0, "dual_moon", packed, k, steps)
        out.update({f"{kind}_key": nk, f"{kind}_pos": pos, f"{kind}_lp": lp, f"{kind}_acc": acc})
    np.savez_compressed(os.path.join(OUT, "local_dualmoon_d5.npz"), key=ks[0], x0=x0, **out)


def nf_case():
    d, n, n_steps = 5, 6, 7
    p = random_params(21, d, 2, [16, 16], 8, gain=1.0, affine=0.0, whiten=False)
    key = rng.PRNGKey(7)
    ks = rng.split(key, 2)
    x0 = rng.normal(ks[1], (n, d))
    packed = otargets.IsoGaussian.pack(d, 0.5)
    out = {}
    for tag, bs in (("simple", 100), ("batched", 3)):
        nk, pos, lp, acc, last, dbg = nf.take_group_steps(ks[0], x0, p, "iso_gaussian", packed, n_steps, bs)
        out.update({f"{tag}_key": nk, f"{tag}_pos": pos, f"{tag}_lp": lp, f"{tag}_acc": acc,
                    f"{tag}_proposals": dbg["proposals"], f"{tag}_lp_nf": dbg["lp_nf_prop"]})
    np.savez_compressed(os.path.join(OUT, "nf_global_iso_d5.npz"), key=ks[0], x0=x0, params_flat=nf.flatten(p), **out)


def rng_case():
    key = rng.PRNGKey(42)
    np.savez_compressed(os.path.join(OUT, "rng_key42.npz"), split=rng.split(key, 4), bits=rng.random_bits(key, (16,)),
                        uniform=rng.uniform(key, (16,)), normal=rng.normal(key, (16,)),
                        permutation=rng.permutation(key, 2000), choice=rng.choice_with_replacement(key, 1000, 64))


def strategies_case():
    """AdamOptimization and ParallelTempering on the reference's own test set-up
    (test/unit/test_strategies.py:27-78, 337-392)."""
    key = rng.PRNGKey(42)
    key, sub = rng.split(key)
    x0 = (rng.normal(sub, (20, 2)) * 1 + 10).astype(np.float32)
    data2 = otargets.IsoGaussian.pack(2, 0.5, np.arange(2))
    a_key, a_x, a_lp = oopt.adam_optimize(key, "iso_gaussian", data2, x0, n_steps=100, learning_rate=5e-2, noise_level=0.0)
    n_key, n_x, n_lp = oopt.adam_optimize(key, "iso_gaussian", data2, x0, n_steps=30, learning_rate=1e-2, noise_level=10.0,
                                          bounds=[[9.0, 10.5]])
    key = rng.PRNGKey(42)
    key, sub = rng.split(key)
    p0 = rng.normal(sub, (7, 3))
    key, sub = rng.split(key)
    tp = rng.normal(sub, (7, 4, 3))
    temps = (np.arange(5) + 1.0).astype(np.float32)
    data3 = otargets.IsoGaussian.pack(3, 0.5, np.arange(3))
    pt_key, pt_p0, pt_tp, pt_t, pt_acc = opt_pt.parallel_tempering(key, p0, tp, temps, "iso_gaussian", data3, 4, 1.0)
    np.savez_compressed(os.path.join(OUT, "strategies_iso.npz"), adam_key=a_key, adam_x0=x0, adam_x=a_x, adam_lp=a_lp,
                        adam_noisy_key=n_key, adam_noisy_x=n_x, adam_noisy_lp=n_lp, pt_key_in=key, pt_x0=p0,
                        pt_tempered_in=tp, pt_key=pt_key, pt_positions=pt_p0, pt_tempered=pt_tp,
                        pt_temperatures=pt_t, pt_accepts=pt_acc)


def realnvp_params(seed, d, L, h, gain=30.0):
    """RealNVP parameters away from the near-identity initialisation (first-layer weights are drawn with
    std sqrt(1e-4 / d)): scaled-up first layers, non-trivial whitening constants, and masks part-way through the
    weight-decay shrinkage the reference applies to them (1 -> 0.97)."""
    r = np.random.default_rng(seed)
    p = onvp.init_params(rng.PRNGKey(seed), d, L, h)
    p.W1 = (p.W1 * np.float32(gain)).astype(np.float32)
    p.mask = (p.mask * np.float32(0.97)).astype(np.float32)
    p.data_mean = r.standard_normal(d).astype(np.float32)
    a = r.standard_normal((d, d)) * 0.3
    p.data_cov = (a @ a.T + np.diag(0.5 + r.random(d))).astype(np.float32)
    p.base_cov = (p.base_cov * np.float32(0.98)).astype(np.float32)
    return p


def realnvp_case():
    """RealNVP: bijection, log_prob, sample, loss gradient, bit-level initialisation and two
    NFProposal runs, on the shapes of test/unit/test_nf.py:24-52 and a wider one."""
    d, L, h, n = 5, 4, 16, 24
    p = realnvp_params(31, d, L, h)
    x = (2.0 * np.random.default_rng(3).standard_normal((n, d))).astype(np.float32)
    y, ld = onvp.forward(p, x)
    xi, ldi = onvp.inverse(p, x)
    key = rng.PRNGKey(123)
    loss, g = onvp.loss_and_grads(p, x)
    blob = dict(x=x, fwd_y=y, fwd_logdet=ld, inv_x=xi, inv_logdet=ldi, log_prob=onvp.log_prob(p, x), sample_key=key,
                sample=onvp.sample(p, key, 16), params_flat=onvp.flatten(p), shape=np.array([d, L, h], np.int32),
                loss=np.float32(loss), grad_flat=onvp.flatten(g, p))
    p0 = onvp.init_params(rng.PRNGKey(0), 3, 2, 4)          # test_nf.py:30-31 builds RealNVP(3, 2, 4, key)
    blob.update(init_key0_flat=onvp.flatten(p0),
                init_key0_log_prob=onvp.log_prob(p0, np.array([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]], np.float32)))
    ks = rng.split(rng.PRNGKey(7), 2)
    x0 = rng.normal(ks[1], (6, d))
    packed = otargets.IsoGaussian.pack(d, 0.5)
    for tag, bs in (("simple", 100), ("batched", 3)):
        nk, pos, lp, acc, last, dbg = nf.take_group_steps(ks[0], x0, p, "iso_gaussian", packed, 7, bs)
        blob.update({f"nf_{tag}_key": nk, f"nf_{tag}_pos": pos, f"nf_{tag}_lp": lp, f"nf_{tag}_acc": acc,
                     f"nf_{tag}_proposals": dbg["proposals"], f"nf_{tag}_lp_nf": dbg["lp_nf_prop"]})
    blob.update(nf_key=ks[0], nf_x0=x0)
    np.savez_compressed(os.path.join(OUT, "realnvp_d5.npz"), **blob)


if __name__ == "__main__":
    which = sys.argv[1:] or ["flow", "init", "local", "nf", "rng", "strategies", "realnvp"]
    for name in which:
        {"flow": flow_case, "init": init_case, "local": local_case, "nf": nf_case, "rng": rng_case,
         "strategies": strategies_case, "realnvp": realnvp_case}[name]()
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))
