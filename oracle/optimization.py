"""CPU restatement (TEST INFRASTRUCTURE ONLY) of AdamOptimization.optimize,
src/flowMC/strategy/optimization.py:85-164, with optax.adam's update chain (optax 0.2.4, as recalled: scale_by_adam
with bias correction 1 - decay**count, then scale(-learning_rate)) and optax.projections.projection_box = clip.
numpy float32, vectorised over chains.  Parity unpinned against the reference's own outputs (jax / optax are not
installable here); pinned to the invariants the reference tests assert (tests/test_oracle_golden.py)."""
import numpy as np

from . import rng, targets

F32 = np.float32


def adam_optimize(rng_key, target, data, initial_position, n_steps=100, learning_rate=1e-2, noise_level=10.0,
                  bounds=((-np.inf, np.inf),), chain_offset=0, n_chains_total=None, b1=0.9, b2=0.999, eps=1e-8):
    """Returns (new_rng_key, optimized_positions [n, d], final_log_prob [n])."""
    x = np.array(initial_position, dtype=F32)
    n, d = x.shape
    n_tot = n if n_chains_total is None else n_chains_total
    bounds = np.broadcast_to(np.asarray(bounds, dtype=F32), (d, 2))
    lo, hi = bounds[:, 0], bounds[:, 1]
    ks = rng.split(rng_key, 2)                                        # optimization.py:149
    new_key, subkey = ks[0], ks[1]
    keys = rng.split(subkey, n_tot)[chain_offset:chain_offset + n]    # :150
    mu = np.zeros_like(x)
    nu = np.zeros_like(x)
    for t in range(1, int(n_steps) + 1):
        z = np.empty(n, dtype=F32)
        for c in range(n):                                            # :122-123  key, subkey = split(key)
            kk = rng.split(keys[c], 2)
            keys[c] = kk[0]
            z[c] = rng.normal(kk[1], ())                              # scalar draw: counter 0
        _, g = targets.logp_grad(target, x, data)
        s = (F32(1) + z * F32(noise_level)).astype(F32)               # :125-127
        g = ((-g).astype(F32) * s[:, None]).astype(F32)               # grad of -logpdf, noisy
        mu = (F32(1 - b1) * g + F32(b1) * mu).astype(F32)             # optax.scale_by_adam
        nu = (F32(1 - b2) * (g * g) + F32(b2) * nu).astype(F32)
        bc1 = F32(1) - F32(b1) ** F32(t)
        bc2 = F32(1) - F32(b2) ** F32(t)
        u = ((mu / bc1) / (np.sqrt((nu / bc2).astype(F32)) + F32(eps))).astype(F32)
        x = (x + (F32(-learning_rate) * u).astype(F32)).astype(F32)   # scale(-lr), apply_updates
        x = np.minimum(np.maximum(x, lo), hi).astype(F32)             # projection_box (:131-133)
    return new_key, x, targets.logp(target, x, data).astype(F32)
