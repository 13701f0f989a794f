"""NVTX ranges around the host-side phases of the hot path (the reference's only tracing is its
trace-time ``print("Compiling ...")`` lines; the B200 build adds one named range per strategy call and per training
epoch, so that an Nsight Systems / ncu timeline reads as local_stepper / model_trainer / global_stepper).

``FLOWMC_NVTX=0`` switches the ranges off; they cost two library calls each and nothing is recorded unless a profiler
is attached.  The device-side timeline of the tensor-core kernels is ``flowmc_trace_tc_timeline`` (include/flowmc_b200.h).
"""
from __future__ import annotations

import contextlib
import os

_ENABLED = os.environ.get("FLOWMC_NVTX", "1") != "0"
try:
    import torch.cuda.nvtx as _nvtx
except Exception:  # pragma: no cover
    _nvtx = None

ranges_opened = 0   # for tests: how many ranges this process has pushed


@contextlib.contextmanager
def nvtx_range(name: str):
    global ranges_opened
    pushed = False
    if _ENABLED and _nvtx is not None:
        try:
            _nvtx.range_push(name)
            pushed = True
            ranges_opened += 1
        except Exception:   # NVTX library not loadable (CPU-only wheel): tracing is best-effort
            pushed = False
    try:
        yield
    finally:
        if pushed:
            _nvtx.range_pop()
