"""ctypes loader of the C restatement (oracle/c/flowmc_ref.c).  TEST / CPU-BASELINE ONLY.

The reference (flowMC 0.4.5) is Python over jax; jax is not installable in this image, so
"oracle/_ref built from the reference's own sources" does not exist.  This C port is the
multi-threaded CPU stand-in that bench.py times (cpu_baseline.kind == "port").
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent


def _cpu_tag() -> str:
    """The library is built with -march=native and travels with the repo snapshot to the GPU box, whose host CPU may
    differ: one build per CPU flag set."""
    import hashlib
    try:
        flags = next(l for l in open("/proc/cpuinfo") if l.startswith("flags"))
    except Exception:
        flags = "unknown"
    return hashlib.sha1(flags.encode()).hexdigest()[:10]


_SO = _HERE / "_build" / f"libflowmc_ref_{_cpu_tag()}.so"
TARGET_IDS = {"iso_gaussian": 0, "dual_moon": 1, "ar1_gaussian": 2, "dense_gaussian": 3, "rosenbrock": 4,
              "gaussian_mixture": 5}
KIND_IDS = {"MALA": 0, "HMC": 1, "GRW": 2}


def build(force: bool = False) -> Path:
    src = _HERE / "c" / "flowmc_ref.c"
    if force or not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE / "c"), "-B", f"OUT={_SO}"], check=True, capture_output=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_SO))
        _lib.ref_num_threads.restype = C.c_int
        _lib.ref_take_serial_steps.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def num_threads() -> int:
    return int(lib().ref_num_threads())


def set_num_threads(n: int) -> int:
    """OpenMP threads for the following calls (torchrun exports OMP_NUM_THREADS=1); returns the count in effect."""
    lib().ref_set_num_threads(C.c_int(int(n)))
    return num_threads()


def threefry2x32(k0, k1, c0, c1):
    out = np.zeros(2, np.uint32)
    lib().ref_threefry2x32(C.c_uint32(k0), C.c_uint32(k1), C.c_uint32(c0), C.c_uint32(c1), _p(out))
    return int(out[0]), int(out[1])


def normal(key, n):
    key = np.ascontiguousarray(key, np.uint32)
    out = np.empty(n, np.float32)
    lib().ref_normal(_p(key), C.c_int(n), _p(out))
    return out


def take_serial_steps(rng_key, x0, target, data, kind, n_steps, thinning=1, chain_offset=0, step_size=0.1,
                      n_leapfrog=0, chol=None, colsum=None, store=True, debug=False):
    """take_steps.py:60-144 for n chains.  Returns (key_out, positions, log_probs, accepts, last) and, with
    ``debug``, additionally (ratio, log_u) [n, n_out]: the two sides of every stored step's accept test."""
    x0 = np.ascontiguousarray(x0, np.float32)
    n, d = x0.shape
    data = np.ascontiguousarray(data, np.float32)
    key = np.ascontiguousarray(rng_key, np.uint32)
    key_out = np.zeros(2, np.uint32)
    n_out = len(range(0, n_steps, thinning))
    pos = np.empty((n, n_out, d), np.float32) if store else None
    lp = np.empty((n, n_out), np.float32)
    acc = np.empty((n, n_out), np.float32)
    last = np.empty((n, d), np.float32)
    ratio = np.empty((n, n_out), np.float32) if debug else None
    logu = np.empty((n, n_out), np.float32) if debug else None
    chol = None if chol is None else np.ascontiguousarray(chol, np.float32)
    colsum = None if colsum is None else np.ascontiguousarray(colsum, np.float32)
    rc = lib().ref_take_serial_steps(C.c_int(KIND_IDS[kind]), C.c_int(TARGET_IDS[target]), _p(data), _p(key), _p(x0),
                                     C.c_int64(n), C.c_int(d), C.c_int(n_steps), C.c_int(thinning),
                                     C.c_int64(chain_offset), C.c_float(step_size), C.c_int(n_leapfrog), _p(chol),
                                     _p(colsum), _p(key_out), _p(pos), _p(lp), _p(acc), _p(last), _p(ratio), _p(logu))
    if rc != 0:
        raise RuntimeError(f"ref_take_serial_steps failed: {rc}")
    if debug:
        return key_out, pos, lp, acc, last, ratio, logu
    return key_out, pos, lp, acc, last
