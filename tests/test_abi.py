"""The C-ABI library loads without a GPU and exports every symbol include/flowmc_b200.h declares."""
import ctypes
import re
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "flowmc_b200.h").read_text()
    return sorted(set(re.findall(r"FLOWMC_API[^;(]*?\b(flowmc_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    from flowmc_b200._lib import lib
    syms = _declared_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in flowmc_b200.h but not exported"


def test_registry_has_builtin_targets():
    from flowmc_b200._lib import lib
    names = {lib.flowmc_target_name(i).decode() for i in range(lib.flowmc_target_count())}
    assert {"iso_gaussian", "dual_moon", "ar1_gaussian", "dense_gaussian", "rosenbrock", "gaussian_mixture"} <= names
    assert lib.flowmc_target_lookup(b"no_such_target") < 0
    assert b"no_such_target" in lib.flowmc_last_error()


def test_host_key_split_matches_oracle():
    from flowmc_b200 import random as frandom
    from oracle import rng
    for seed in (0, 1, 42, 2**31 + 5):
        k = frandom.PRNGKey(seed)
        assert np.array_equal(k, rng.PRNGKey(seed))
        assert np.array_equal(frandom.split(k, 5), rng.split(k, 5))


def test_python_callable_is_rejected():
    import pytest
    from flowmc_b200.resource.logPDF import LogPDF
    with pytest.raises(TypeError):
        LogPDF(lambda x, data: -0.5 * (x ** 2).sum(), n_dims=3)


def test_user_plugin_compiles_and_registers(tmp_path):
    """A target written against include/flowmc_target.cuh builds into its own .so and registers
    itself on load (no GPU needed: nvcc cross-compiles, registration is a host-side table)."""
    from test_gpu_targets import QUARTIC_PLUGIN
    from flowmc_b200 import targets as T
    from flowmc_b200._lib import lib
    tgt = T.compile_target(QUARTIC_PLUGIN, "test_quartic", lambda data, d: np.array([0.25], np.float32),
                           str(tmp_path))
    assert lib.flowmc_target_lookup(b"test_quartic") == tgt.target_id >= 6
