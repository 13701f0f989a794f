"""CPU restatement (TEST INFRASTRUCTURE ONLY) of ParallelTempering, src/flowMC/strategy/parallel_tempering.py:47-436,
and TemperedPDF.tempered_log_pdf, src/flowMC/resource/logPDF.py:104-106, with MALA (what the reference's PT bundle
uses), HMC or the Gaussian random walk as the tempered kernel (the reference passes any ProposalBase, :74,135-289).
numpy float32; the (chain, temperature) pairs are flattened into rows.  Parity unpinned against the reference's own
outputs (jax is not installable here); the invariants the reference tests assert are checked in
tests/test_oracle_golden.py."""
import numpy as np

from . import local, rng, targets

F32 = np.float32


def log_prior(prior, x):
    """prior = None (flat 0) or [4, d] = c, m, lo, hi of -sum c (x - m)^2 inside the box, -inf outside."""
    x = np.asarray(x, F32)
    if prior is None:
        return np.zeros(x.shape[:-1], F32), np.zeros_like(x)
    c, m, lo, hi = [np.asarray(v, F32) for v in prior]
    r = (x - m).astype(F32)
    val = (-np.sum(c * r * r, axis=-1, dtype=F32)).astype(F32)
    inside = np.all((x >= lo) & (x <= hi), axis=-1)
    return np.where(inside, val, F32(-np.inf)).astype(F32), (F32(-2.0) * c * r).astype(F32)


class _Tempered:
    """Row-wise tempered target: (1 / T_row) * logpdf + log_prior, and its gradient."""
    name = "__tempered__"

    def __init__(self, target, beta, prior):
        self.target, self.beta, self.prior = target, np.asarray(beta, F32), prior

    def logp_grad(self, x, data):
        lp, g = targets.logp_grad(self.target, x, data)
        pl, pg = log_prior(self.prior, x)
        return ((self.beta * lp).astype(F32) + pl).astype(F32), ((self.beta[:, None] * g).astype(F32) + pg).astype(F32)


def ensemble_steps(subkey, positions, target, data, temperatures, n_steps, step_size, prior=None,
                   chain_offset=0, n_chains_total=None, kind="MALA", **kernel_kw):
    """_ensemble_step vmapped over chains (:91-101, :250-290).  positions [n, n_temps, d].  Returns final positions,
    final TEMPERED log-probs [n, n_temps], accept flags [n, n_temps, n_steps]."""
    positions = np.asarray(positions, F32)
    n, n_temps, d = positions.shape
    n_tot = n if n_chains_total is None else n_chains_total
    chain_keys = rng.split(subkey, n_tot)[chain_offset:chain_offset + n]           # :98
    keys = np.stack([rng.split(k, n_temps) for k in chain_keys]).reshape(n * n_temps, 2)   # :283
    beta = np.tile((F32(1.0) / np.asarray(temperatures, F32)).astype(F32), n)      # logPDF.py:106
    tgt = _Tempered(target, beta, prior)
    local.TARGETS[tgt.name] = tgt
    kernel = local.make_kernel(kind, step_size=step_size, **kernel_kw)
    try:
        x = positions.reshape(n * n_temps, d).copy()
        lp, _ = tgt.logp_grad(x, data)                                             # :229-231
        accs = np.zeros((n * n_temps, n_steps), F32)
        for t in range(n_steps):                                                   # _individual_step_body (:189-197)
            s = rng.split(keys, 2)
            keys, sub = s[:, 0, :], s[:, 1, :]
            x, lp, acc, _ = kernel(sub, x, lp, tgt.name, data)
            accs[:, t] = acc
    finally:
        del local.TARGETS[tgt.name]
    return x.reshape(n, n_temps, d), lp.reshape(n, n_temps), accs.reshape(n, n_temps, n_steps)


def exchange(subkey, positions, target, data, temperatures, chain_offset=0, n_chains_total=None):
    """_exchange vmapped over chains (:110-115, :291-398).  Returns positions, UNtempered log-probs (both after the
    swaps), accept flags [n, n_temps - 1], and the (ratio, log_uniform) pairs for near-tie analysis."""
    positions = np.array(positions, F32)
    n, n_temps, d = positions.shape
    n_tot = n if n_chains_total is None else n_chains_total
    T = np.asarray(temperatures, F32)
    keys = rng.split(subkey, n_tot)[chain_offset:chain_offset + n]
    lp = targets.logp(target, positions.reshape(n * n_temps, d), data).reshape(n, n_temps).astype(F32)   # :381
    accs = np.zeros((n, n_temps - 1), F32)
    ratios = np.zeros((n, n_temps - 1), F32)
    logus = np.zeros((n, n_temps - 1), F32)
    for idx in range(n_temps - 1):
        s = rng.split(keys, 2)
        keys, sub = s[:, 0, :], s[:, 1, :]
        ratio = ((F32(1.0) / T[idx + 1] - F32(1.0) / T[idx]).astype(F32) * (lp[:, idx] - lp[:, idx + 1]).astype(F32)).astype(F32)
        with np.errstate(divide="ignore"):
            log_uniform = np.log(rng.uniform(sub, ())).astype(F32)
        acc = log_uniform < ratio
        for c in np.nonzero(acc)[0]:
            positions[c, [idx, idx + 1]] = positions[c, [idx + 1, idx]]
            lp[c, [idx, idx + 1]] = lp[c, [idx + 1, idx]]
        accs[:, idx], ratios[:, idx], logus[:, idx] = acc, ratio, log_uniform
    return positions, lp, accs, ratios, logus


def adapt_temperature(temperatures, do_accept):
    """_adapt_temperature (:400-436) in float32."""
    t = np.asarray(temperatures, F32)
    acc = np.asarray(do_accept, F32)
    rate = acc.mean(axis=0, dtype=F32)
    damping = (F32(100.0 / acc.shape[0]) * (rate[:-1] - rate[1:])).astype(F32)
    new_t = t.copy()
    for i in range(1, t.shape[0] - 1):
        new_t[i] = new_t[i - 1] + (t[i] - t[i - 1]) * np.exp(damping[i - 1], dtype=F32)
    return new_t


def parallel_tempering(rng_key, initial_position, tempered_positions, temperatures, target, data, n_steps, step_size,
                       prior=None, training=True, kind="MALA", **kernel_kw):
    """ParallelTempering.__call__ (:47-132).  Returns (rng_key, positions[:, 0], new tempered positions, new
    temperatures, exchange accepts)."""
    rng_key, _ = rng.split(rng_key)                                                # :73
    positions = np.concatenate([np.asarray(initial_position, F32)[:, None, :], np.asarray(tempered_positions, F32)], axis=1)
    rng_key, subkey = rng.split(rng_key)
    positions, _, _ = ensemble_steps(subkey, positions, target, data, temperatures, n_steps, step_size, prior,
                                     kind=kind, **kernel_kw)
    rng_key, subkey = rng.split(rng_key)
    positions, _, accs, _, _ = exchange(subkey, positions, target, data, temperatures)
    temps = np.asarray(temperatures, F32)
    tempered = np.asarray(tempered_positions, F32)
    if training:
        tempered = positions[:, 1:]
        temps = adapt_temperature(temps, accs)
    return rng_key, positions[:, 0], tempered, temps, accs
