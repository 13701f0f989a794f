"""Oracle: jax.random semantics (threefry2x32, partitionable layout) in numpy.  TEST ONLY.

The reference draws every random number through ``jax.random`` (call sites:
src/flowMC/strategy/take_steps.py:71-72,158; resource/kernel/MALA.py:62,66,83;
HMC.py:128-136,144; Gaussian_random_walk.py:48-56; NF_proposal.py:41,99,103;
strategy/train_model.py:72-81; resource/model/nf_model/base.py:141,191;
resource/model/common.py:93-107,291-293).  jax itself is a third-party dependency absent
from /root/reference (pinned jax==0.5.0 / jaxlib==0.5.0 in uv.lock:786-787,847-848), so
this file restates its *published* algorithm:

  * threefry2x32-20 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11),
  * ``jax_threefry_partitionable=True`` (the default from jax 0.5.0): element with row-major
    flat index i uses counter (hi32(i), lo32(i)); ``split`` returns both output words,
    ``random_bits(32)`` returns their XOR,
  * ``uniform``: mantissa-fill of the top 23 bits, ``normal``: sqrt(2)*erf_inv(uniform(-1,1)),
    with XLA's float32 erf_inv polynomial (Giles 2010 single-precision approximation).

Pinned by tests/test_oracle_rng.py (Random123 KATs, documented split(PRNGKey(42)) values,
scipy erfinv).  Everything here is uint32 / float32 numpy; no float64 leaks into results.
"""
from __future__ import annotations

import math

import numpy as np

U32 = np.uint32
F32 = np.float32

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_PARITY = U32(0x1BD11BDA)


def PRNGKey(seed: int) -> np.ndarray:
    """jax.random.PRNGKey(seed) with x64 disabled: uint32[2] = [0, seed]."""
    return np.array([0, seed & 0xFFFFFFFF], dtype=U32)


def _rotl(x: np.ndarray, r: int) -> np.ndarray:
    return (x << U32(r)) | (x >> U32(32 - r))


def threefry2x32(k0, k1, c0, c1):
    """threefry2x32-20 block function on broadcastable uint32 arrays -> (o0, o1)."""
    k0 = np.asarray(k0, dtype=U32)
    k1 = np.asarray(k1, dtype=U32)
    c0 = np.asarray(c0, dtype=U32)
    c1 = np.asarray(c1, dtype=U32)
    with np.errstate(over="ignore"):
        ks = (k0, k1, k0 ^ k1 ^ _PARITY)
        x0 = c0 + ks[0]
        x1 = c1 + ks[1]
        for g in range(5):
            for r in _ROT[g % 2]:
                x0 = x0 + x1
                x1 = _rotl(x1, r)
                x1 = x1 ^ x0
            x0 = x0 + ks[(g + 1) % 3]
            x1 = x1 + ks[(g + 2) % 3] + U32(g + 1)
    return x0.astype(U32), x1.astype(U32)


def _counters(n: int):
    idx = np.arange(n, dtype=np.uint64)
    return (idx >> np.uint64(32)).astype(U32), (idx & np.uint64(0xFFFFFFFF)).astype(U32)


def split(key: np.ndarray, num: int = 2) -> np.ndarray:
    """jax.random.split(key, num) -> uint32[num, 2].  ``key`` may be batched: [..., 2]."""
    key = np.asarray(key, dtype=U32)
    hi, lo = _counters(num)
    o0, o1 = threefry2x32(key[..., 0:1], key[..., 1:2], hi, lo)
    return np.stack([o0, o1], axis=-1)


def random_bits(key: np.ndarray, shape) -> np.ndarray:
    """jax.random.bits(key, shape, uint32).  ``key`` may be batched: [..., 2] -> [..., *shape]."""
    key = np.asarray(key, dtype=U32)
    shape = tuple(shape) if not isinstance(shape, int) else (shape,)
    n = int(np.prod(shape)) if shape else 1
    hi, lo = _counters(n)
    o0, o1 = threefry2x32(key[..., 0:1], key[..., 1:2], hi, lo)
    bits = o0 ^ o1
    return bits.reshape(key.shape[:-1] + shape)


def bits_to_unit_float(bits: np.ndarray) -> np.ndarray:
    """[0,1) float32 from 32 random bits: bitcast((bits >> 9) | 0x3F800000) - 1."""
    fb = (bits >> U32(9)) | U32(0x3F800000)
    return fb.view(F32) - F32(1.0)


def uniform(key, shape=(), minval=0.0, maxval=1.0) -> np.ndarray:
    bits = random_bits(key, shape)
    f = bits_to_unit_float(bits)
    lo = F32(minval)
    hi = F32(maxval)
    return np.maximum(lo, f * F32(hi - lo) + lo).astype(F32)


_ERFINV_LT5 = [2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06,
               0.00021858087, -0.00125372503, -0.00417768164, 0.246640727, 1.50140941]
_ERFINV_GE5 = [-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844,
               0.00573950773, -0.0076224613, 0.00943887047, 1.00167406, 2.83297682]


def erf_inv32(x: np.ndarray) -> np.ndarray:
    """XLA's float32 erf_inv (xla/client/lib/math.cc ErfInv32), all ops in float32."""
    x = np.asarray(x, dtype=F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = -np.log1p(-(x * x)).astype(F32)
        lt = w < F32(5.0)
        w2 = np.where(lt, w - F32(2.5), np.sqrt(np.maximum(w, F32(0))).astype(F32) - F32(3.0)).astype(F32)
        p = np.where(lt, F32(_ERFINV_LT5[0]), F32(_ERFINV_GE5[0])).astype(F32)
        for a, b in zip(_ERFINV_LT5[1:], _ERFINV_GE5[1:]):
            c = np.where(lt, F32(a), F32(b)).astype(F32)
            p = (c + p * w2).astype(F32)
        res = (p * x).astype(F32)
        res = np.where(np.abs(x) == F32(1.0), x * F32(np.inf), res)
    return res.astype(F32)


_NORMAL_LO = np.nextafter(F32(-1.0), F32(0.0), dtype=F32)
_SQRT2 = F32(np.sqrt(2))


def bits_to_normal(bits: np.ndarray) -> np.ndarray:
    f = bits_to_unit_float(bits)
    span = F32(F32(1.0) - _NORMAL_LO)  # rounds to 2.0f
    u = np.maximum(_NORMAL_LO, (f * span).astype(F32) + _NORMAL_LO).astype(F32)
    return (_SQRT2 * erf_inv32(u)).astype(F32)


def normal(key, shape=()) -> np.ndarray:
    return bits_to_normal(random_bits(key, shape))


def randint(key, shape, minval: int, maxval: int) -> np.ndarray:
    """jax.random.randint for int32 (double-width rejection-free scheme, biased as in jax)."""
    k = split(key, 2)
    hi = random_bits(k[0], shape)
    lo = random_bits(k[1], shape)
    span = U32(max(1, maxval - minval))
    with np.errstate(over="ignore"):
        mult = U32(U32(1 << 16) % span)
        mult = U32(U32(mult * mult) % span)
        off = (hi % span) * mult + (lo % span)
        off = off % span
    return (np.int64(minval) + off.astype(np.int64)).astype(np.int32)


def choice_with_replacement(key, n: int, m: int) -> np.ndarray:
    """jax.random.choice(key, arange(n), (m,), replace=True) == randint(key, (m,), 0, n)."""
    return randint(key, (m,), 0, n)


def permutation(key, n: int) -> np.ndarray:
    """jax.random.permutation(key, n): repeated stable sort by random 32-bit keys."""
    x = np.arange(n, dtype=np.int32)
    rounds = int(np.ceil(3 * np.log(max(1, n)) / np.log(np.iinfo(np.uint32).max)))
    key = np.asarray(key, dtype=U32)
    for _ in range(rounds):
        ks = split(key, 2)
        key, sub = ks[0], ks[1]
        sk = random_bits(sub, (n,))
        order = np.argsort(sk, kind="stable")
        x = x[order]
    return x


def multivariate_normal_identity(key, n: int, d: int) -> np.ndarray:
    """jax.random.multivariate_normal(key, 0, I, (n,)) = normal(key, (n, d)) @ chol(I).T"""
    return normal(key, (n, d))
