"""ctypes binding of libflowmc_b200.so (the C ABI declared in include/flowmc_b200.h).

There is deliberately no fallback: if the shared object is missing or a call fails, an
exception is raised.  Nothing here imports the CPU oracle.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libflowmc_b200.so"


class FlowmcError(RuntimeError):
    pass


class LocalParams(C.Structure):
    _fields_ = [
        ("step_size", C.c_float),
        ("n_leapfrog", C.c_int),
        ("hmc_chol", C.c_void_p),
        ("hmc_colsum", C.c_void_p),
        ("hmc_chol_diagonal", C.c_int),
        ("layout_hint", C.c_int),
        ("step_keys", C.c_void_p),
        ("lp0", C.c_void_p),
        ("workspace", C.c_void_p),
        ("workspace_bytes", C.c_int64),
        ("chain_keys", C.c_void_p),
        ("beta", C.c_void_p),
        ("prior", C.c_void_p),
        ("force_n_seg", C.c_int),
        ("slots_override", C.c_int),
    ]


class FlowDesc(C.Structure):
    _fields_ = [
        ("n_features", C.c_int), ("n_layers", C.c_int), ("n_linear", C.c_int), ("num_bins", C.c_int),
        ("dims", C.c_int * 5),
        ("range_min", C.c_float), ("range_max", C.c_float),
        ("off_W", C.c_int64 * 4), ("off_b", C.c_int64 * 4), ("off_scale", C.c_int64), ("off_shift", C.c_int64),
        ("layer_stride", C.c_int64),
        ("off_data_mean", C.c_int64), ("off_data_cov", C.c_int64), ("off_base_mean", C.c_int64),
        ("off_base_cov", C.c_int64),
        ("n_params", C.c_int64),
        ("tc_image", C.c_void_p),
        ("tc_terms", C.c_int),
    ]


class RealNVPDesc(C.Structure):
    _fields_ = [
        ("n_features", C.c_int), ("n_layers", C.c_int), ("n_hidden", C.c_int), ("dt", C.c_float),
        ("off_W1s", C.c_int64), ("off_b1s", C.c_int64), ("off_W2s", C.c_int64), ("off_b2s", C.c_int64),
        ("off_W1t", C.c_int64), ("off_b1t", C.c_int64), ("off_W2t", C.c_int64), ("off_b2t", C.c_int64),
        ("off_mask", C.c_int64), ("layer_stride", C.c_int64),
        ("off_data_mean", C.c_int64), ("off_data_cov", C.c_int64), ("off_base_mean", C.c_int64),
        ("off_base_cov", C.c_int64), ("n_params", C.c_int64),
    ]


class GlobalParams(C.Structure):
    _fields_ = [
        ("n_batch_size", C.c_int),
        ("chain_keys", C.c_void_p),
        ("lp0", C.c_void_p),
        ("workspace", C.c_void_p),
        ("workspace_bytes", C.c_int64),
    ]


def _load() -> C.CDLL:
    if not _LIB_PATH.exists():
        raise ImportError(
            f"{_LIB_PATH} not found: the CUDA library has not been built. "
            "Run `python -m flowmc_b200.build` (needs nvcc); there is no CPU fallback."
        )
    lib = C.CDLL(str(_LIB_PATH), mode=C.RTLD_GLOBAL)
    vp, i32, i64, f32, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double
    u32p = C.POINTER(C.c_uint32)
    sigs = {
        "flowmc_abi_version": (i32, []),
        "flowmc_last_error": (C.c_char_p, []),
        "flowmc_target_count": (i32, []),
        "flowmc_target_lookup": (i32, [C.c_char_p]),
        "flowmc_target_name": (C.c_char_p, [i32]),
        "flowmc_target_eval": (i32, [i32, vp, vp, i64, i32, vp, vp, vp]),
        "flowmc_key_split": (i32, [u32p, i64, u32p]),
        "flowmc_key_split_batch": (i32, [u32p, i64, i64, u32p]),
        "flowmc_random_bits": (i32, [u32p, i64, vp, vp]),
        "flowmc_random_uniform": (i32, [u32p, i64, f32, f32, vp, vp]),
        "flowmc_random_normal": (i32, [u32p, i64, vp, vp]),
        "flowmc_local_steps": (i32, [i32, i32, vp, u32p, vp, vp, vp, vp, i64, i64, i64, i32, i32, i32, i64, i64,
                                     C.POINTER(LocalParams), u32p, vp, vp]),
        "flowmc_local_steps_plan": (i32, [i32, i32, i64, i32, i32, C.POINTER(LocalParams), C.POINTER(C.c_int)]),
        "flowmc_adam_optimize": (i32, [i32, vp, u32p, vp, i64, i32, i32, f32, f32, vp, vp, vp, i64, i64, u32p, vp, vp,
                                       vp]),
        "flowmc_pt_exchange": (i32, [u32p, i64, i64, i64, i32, i32, vp, vp, vp, vp, vp]),
        "flowmc_launch_count": (i64, []),
        "flowmc_local_steps_workspace_bytes": (i64, [i64, i32, i32]),
        "flowmc_flow_desc_init": (i32, [C.POINTER(FlowDesc), i32, i32, i32, C.POINTER(C.c_int), i32, f32, f32]),
        "flowmc_flow_tc_image_bytes": (i64, [C.POINTER(FlowDesc)]),
        "flowmc_flow_tc_pack": (i32, [C.POINTER(FlowDesc), vp, vp, vp]),
        "flowmc_flow_forward": (i32, [C.POINTER(FlowDesc), vp, vp, i64, vp, vp, vp]),
        "flowmc_flow_inverse": (i32, [C.POINTER(FlowDesc), vp, vp, i64, vp, vp, vp]),
        "flowmc_flow_log_prob": (i32, [C.POINTER(FlowDesc), vp, vp, i64, vp, vp, vp]),
        "flowmc_flow_sample": (i32, [C.POINTER(FlowDesc), vp, vp, u32p, i64, i64, vp, vp]),
        "flowmc_flow_loss_grad_workspace_bytes": (i64, [C.POINTER(FlowDesc), i64]),
        "flowmc_flow_loss_grad": (i32, [C.POINTER(FlowDesc), vp, vp, vp, i64, f32, vp, vp, vp, i64, vp]),
        "flowmc_clip_adamw": (i32, [i64, vp, vp, vp, vp, i64, f64, f64, f64, f64, f64, f64, vp, vp, vp]),
        "flowmc_random_permutation_workspace_bytes": (i64, [i64]),
        "flowmc_random_permutation": (i32, [u32p, i64, vp, vp, i64, vp]),
        "flowmc_random_choice": (i32, [u32p, i64, i64, vp, vp]),
        "flowmc_buffer_finite_rows": (i32, [vp, i64, i64, i32, vp, vp, vp, vp]),
        "flowmc_gather_training_rows": (i32, [vp, vp, i64, i32, i32, i32, i64, i64, vp, i64, vp, vp]),
        "flowmc_data_mean_cov": (i32, [vp, i64, i32, vp, vp, vp, vp]),
        "flowmc_trace_tc_timeline": (None, [vp]),
        "flowmc_realnvp_desc_init": (i32, [C.POINTER(RealNVPDesc), i32, i32, i32, f32]),
        "flowmc_realnvp_forward": (i32, [C.POINTER(RealNVPDesc), vp, vp, i64, vp, vp, vp]),
        "flowmc_realnvp_inverse": (i32, [C.POINTER(RealNVPDesc), vp, vp, i64, vp, vp, vp]),
        "flowmc_realnvp_log_prob": (i32, [C.POINTER(RealNVPDesc), vp, vp, i64, vp, vp]),
        "flowmc_realnvp_sample": (i32, [C.POINTER(RealNVPDesc), vp, vp, u32p, i64, i64, vp, vp]),
        "flowmc_realnvp_loss_grad_workspace_bytes": (i64, [C.POINTER(RealNVPDesc), i64]),
        "flowmc_realnvp_loss_grad": (i32, [C.POINTER(RealNVPDesc), vp, vp, vp, i64, f32, vp, vp, vp, i64, vp]),
        "flowmc_nf_accept_scan": (i32, [vp, i64, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i64, vp, vp]),
        "flowmc_peer_block_bytes": (i64, [i64]),
        "flowmc_peer_block_offset": (i64, [i64, i32]),
        "flowmc_ipc_alloc": (i32, [i64, C.POINTER(vp), C.c_char_p]),
        "flowmc_ipc_open": (i32, [C.c_char_p, C.POINTER(vp)]),
        "flowmc_ipc_close": (i32, [vp]),
        "flowmc_ipc_free": (i32, [vp]),
        "flowmc_dp_reduce_adamw": (i32, [i32, i32, C.POINTER(vp), i64, vp, vp, vp, i64, f64, f64, f64, f64, f64, f64,
                                       C.c_uint32, vp, vp]),
        "flowmc_nf_global_steps_workspace_bytes": (i64, [i64, i32, i32]),
        "flowmc_nf_global_steps": (i32, [C.POINTER(FlowDesc), vp, i32, vp, u32p, vp, vp, vp, vp, i64, i64, i64, i32,
                                         i32, i64, i64, C.POINTER(GlobalParams), u32p, vp, vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc: int) -> int:
    if rc < 0:
        raise FlowmcError(f"flowmc_b200 error {rc}: {lib.flowmc_last_error().decode()}")
    return rc


def load_test_lib() -> C.CDLL:
    """libflowmc_b200_test.so: probe kernels for the tcgen05 building blocks (include/flowmc_b200_test.h) -- test
    scaffolding kept out of the product library."""
    path = _LIB_PATH.parent / "libflowmc_b200_test.so"
    if not path.exists():
        raise ImportError(f"{path} not found: run `python -m flowmc_b200.build`")
    t = C.CDLL(str(path), mode=C.RTLD_GLOBAL)
    vp, i32 = C.c_void_p, C.c_int
    for name in ("flowmc_test_tc_gemm", "flowmc_test_tc_gemm_pair"):
        fn = getattr(t, name)
        fn.restype = i32
        fn.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp]
    return t


def load_plugin(path: str) -> None:
    """dlopen a target plugin built against include/flowmc_target.cuh (registers itself)."""
    C.CDLL(str(path), mode=C.RTLD_GLOBAL)
