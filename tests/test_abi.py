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


def test_xla_ffi_shim_source(tmp_path):
    """flowmc_b200/csrc/flowmc_xla_ffi.cc (the jax.ffi adapter north_star names): without jaxlib's xla/ffi/api/ffi.h it
    must compile to an EMPTY translation unit (what this image builds); against a mock of the XLA FFI surface
    (tests/mock_xla) every handler must type-check against include/flowmc_b200.h and define its handler symbol."""
    import subprocess
    src = ROOT / "flowmc_b200" / "csrc" / "flowmc_xla_ffi.cc"
    inc = ["-I", str(ROOT / "include"), "-I", "/usr/local/cuda/include"]
    empty = tmp_path / "empty.o"
    subprocess.run(["g++", "-std=c++17", "-c", *inc, str(src), "-o", str(empty)], check=True)
    nm = subprocess.run(["nm", str(empty)], capture_output=True, text=True).stdout
    assert "Flowmc" not in nm and "flowmc" not in nm
    mocked = tmp_path / "mocked.o"
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-c", "-I", str(ROOT / "tests" / "mock_xla"), *inc, str(src),
                        "-o", str(mocked)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    nm = subprocess.run(["nm", str(mocked)], capture_output=True, text=True).stdout
    handlers = set(re.findall(r"\b(Flowmc[A-Za-z]+)_mock_symbol", nm))
    assert {"FlowmcLocalSteps", "FlowmcKernelStep", "FlowmcTargetEval", "FlowmcGlobalSteps", "FlowmcFlowApply",
            "FlowmcFlowSample", "FlowmcFlowTcPack", "FlowmcFlowLossGrad", "FlowmcClipAdamW",
            "FlowmcRandomPermutation", "FlowmcRandomChoice", "FlowmcBufferFiniteRows", "FlowmcGatherTrainingRows",
            "FlowmcDataMeanCov", "FlowmcPtExchange", "FlowmcAdamOptimize"} <= handlers
    # every C-ABI compute entry point is reachable from some handler
    called = set(re.findall(r"\bU (flowmc_[a-z0-9_]+)", nm))
    for s in ("flowmc_local_steps", "flowmc_nf_global_steps", "flowmc_flow_forward", "flowmc_flow_inverse",
              "flowmc_flow_log_prob", "flowmc_flow_sample", "flowmc_flow_loss_grad", "flowmc_clip_adamw",
              "flowmc_target_eval", "flowmc_pt_exchange", "flowmc_adam_optimize", "flowmc_flow_tc_pack"):
        assert s in called, f"no XLA FFI handler forwards to {s}"


def test_reference_module_spellings():
    """BASELINE.json's north_star spells the modules flowMC.resource.local_kernel.* and flowMC.resource.nf_model.*;
    the checkout has resource.kernel.* and resource.model.nf_model.* (SURVEY 0.4).  Both import the same classes."""
    from flowmc_b200.resource.kernel.MALA import MALA
    from flowmc_b200.resource.kernel.HMC import HMC
    from flowmc_b200.resource.kernel.Gaussian_random_walk import GaussianRandomWalk
    from flowmc_b200.resource.local_kernel.MALA import MALA as MALA2
    from flowmc_b200.resource.local_kernel.HMC import HMC as HMC2
    from flowmc_b200.resource.local_kernel.Gaussian_random_walk import GaussianRandomWalk as GRW2
    from flowmc_b200.resource.local_kernel.base import ProposalBase as PB2
    from flowmc_b200.resource.kernel.base import ProposalBase
    assert MALA is MALA2 and HMC is HMC2 and GaussianRandomWalk is GRW2 and ProposalBase is PB2
    from flowmc_b200.resource.model.nf_model.rqSpline import MaskedCouplingRQSpline
    from flowmc_b200.resource.nf_model.rqSpline import MaskedCouplingRQSpline as M2
    from flowmc_b200.resource.model.nf_model.base import NFModel
    from flowmc_b200.resource.nf_model.base import NFModel as N2
    assert MaskedCouplingRQSpline is M2 and NFModel is N2
