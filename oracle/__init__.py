"""CPU oracle for the flowMC sampling hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (numpy fp32 / uint32, torch-CPU only for autograd
cross-checks) of the reference's algorithm for the path named in BASELINE.json.  It is
imported ONLY by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs, as the *checker* -- never by the product package ``flowmc_b200``.

PARITY STATUS: **parity unpinned at the bit level.**  The reference (flowMC 0.4.5) is pure
Python over jax/equinox/optax, none of which exist in this image (no wheels, no network), so
neither the reference nor JAX could be executed to produce golden vectors.  The reference's
own tests pin no numeric values on this path.  The oracle is therefore pinned against:
  * the three Random123 threefry2x32-20 known-answer vectors,
  * the documented ``jax.random.split(PRNGKey(42))`` values (partitionable threefry, the
    jax>=0.5.0 default pinned by the reference's uv.lock),
  * scipy float64 ``erfinv`` / ``norm`` statistics for the normal sampler,
  * the one KAT in the reference tree: dual-moon(zeros(5)) = -218.14496
    (docs/tutorials/dualmoon.ipynb:65),
  * the reference's invariant tests re-expressed on the oracle (determinism, leapfrog
    reversibility, accept->1 at tiny step, stationarity of mean/var, flow forward/inverse
    consistency), test/unit/test_kernels.py and test/unit/test_nf.py,
  * an independent second restatement in C (oracle/c/) that must agree bit-for-bit on RNG
    words and accept flags.
"""
