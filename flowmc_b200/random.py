"""jax.random look-alikes backed by the library's threefry2x32 (bit-compatible with jax >= 0.5.0).

Keys are host ``numpy.uint32[2]`` arrays (``PRNGKey(seed) == [0, seed]``).  ``split`` runs on
the host inside the C library; ``bits``/``uniform``/``normal`` are CUDA kernels that write
straight into a device tensor.  Used where the reference calls ``jax.random`` outside its
kernels: drawing initial positions in user scripts, strategy key fan-out
(src/flowMC/strategy/take_steps.py:71-72) and model initialisation
(src/flowMC/resource/model/common.py:93-107).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from ._lib import check, lib

_u32p = C.POINTER(C.c_uint32)


def _kp(key: np.ndarray):
    key = np.ascontiguousarray(key, dtype=np.uint32)
    if key.shape != (2,):
        raise ValueError(f"a PRNG key is uint32[2], got shape {key.shape}")
    return key, key.ctypes.data_as(_u32p)


def PRNGKey(seed: int) -> np.ndarray:
    return np.array([0, int(seed) & 0xFFFFFFFF], dtype=np.uint32)


key = PRNGKey


def split_each(keys: np.ndarray, num: int = 2) -> np.ndarray:
    """``jax.vmap(lambda k: jax.random.split(k, num))(keys)``: keys [n, 2] -> [n, num, 2] (one C call)."""
    keys = np.ascontiguousarray(keys, dtype=np.uint32).reshape(-1, 2)
    out = np.empty((keys.shape[0], num, 2), dtype=np.uint32)
    check(lib.flowmc_key_split_batch(keys.ctypes.data_as(_u32p), keys.shape[0], num, out.ctypes.data_as(_u32p)))
    return out


def split(key: np.ndarray, num: int = 2) -> np.ndarray:
    k, kp = _kp(key)
    out = np.empty((num, 2), dtype=np.uint32)
    check(lib.flowmc_key_split(kp, num, out.ctypes.data_as(_u32p)))
    return out


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _device(device):
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def bits(key: np.ndarray, shape=(), device=None) -> torch.Tensor:
    k, kp = _kp(key)
    shape = (shape,) if isinstance(shape, int) else tuple(shape)
    n = math.prod(shape)
    dev = _device(device)
    out = torch.empty(shape, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib.flowmc_random_bits(kp, n, out.data_ptr(), _stream()))
    return out


def uniform(key: np.ndarray, shape=(), minval: float = 0.0, maxval: float = 1.0, device=None) -> torch.Tensor:
    k, kp = _kp(key)
    shape = (shape,) if isinstance(shape, int) else tuple(shape)
    dev = _device(device)
    out = torch.empty(shape, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.flowmc_random_uniform(kp, math.prod(shape), float(minval), float(maxval), out.data_ptr(), _stream()))
    return out


def normal(key: np.ndarray, shape=(), device=None) -> torch.Tensor:
    k, kp = _kp(key)
    shape = (shape,) if isinstance(shape, int) else tuple(shape)
    dev = _device(device)
    out = torch.empty(shape, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.flowmc_random_normal(kp, math.prod(shape), out.data_ptr(), _stream()))
    return out
