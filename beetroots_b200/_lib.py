"""ctypes binding of libbeetroots_b200.so (the C ABI in include/beetroots_b200.h).

There is NO CPU fallback: if the shared library is missing or a call fails, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BT_LIB_PATH", os.path.join(_HERE, "libbeetroots_b200.so"))   # override: A/B runs of two builds

_lib = None

c_double_p = C.POINTER(C.c_double)
c_float_p = C.POINTER(C.c_float)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)
c_uint8_p = C.POINTER(C.c_uint8)

# name -> (restype, argtypes); must list every symbol declared in include/beetroots_b200.h
SIGNATURES = {
    "bt_last_error": (C.c_char_p, []),
    "bt_version": (C.c_int, []),
    "bt_ctx_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "bt_ctx_destroy": (C.c_int, [C.c_void_p]),
    "bt_set_stream": (C.c_int, [C.c_void_p, C.c_uint64]),
    "bt_set_mlp_mode": (C.c_int, [C.c_void_p, C.c_int]),
    "bt_synchronize": (C.c_int, [C.c_void_p]),
    "bt_upload_network": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_int32_p, c_double_p, c_double_p, C.c_int,
                                    c_int32_p, C.POINTER(c_double_p), C.POINTER(c_double_p), c_double_p, c_double_p,
                                    C.c_int]),
    "bt_upload_forward_map": (C.c_int, [C.c_void_p, C.c_int, c_double_p, C.c_int, c_int32_p]),
    "bt_upload_observations": (C.c_int, [C.c_void_p] + [c_double_p] * 6),
    "bt_upload_priors": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p, C.c_int, c_double_p, c_double_p,
                                   C.c_double]),
    "bt_upload_graph": (C.c_int, [C.c_void_p, c_int32_p, C.c_int, c_int32_p, c_int32_p, c_int64_p]),
    "bt_set_state": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_int32_p]),
    "bt_get_state": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_int32_p]),
    "bt_set_theta_f32": (C.c_int, [C.c_void_p, C.c_void_p]),
    "bt_get_theta_f32": (C.c_int, [C.c_void_p, C.c_void_p]),
    "bt_gather_theta_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "bt_scatter_theta_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "bt_compute_all": (C.c_int, [C.c_void_p]),
    "bt_init_v_from_grad": (C.c_int, [C.c_void_p]),
    "bt_get_current": (C.c_int, [C.c_void_p] + [c_double_p] * 6),
    "bt_objective_terms": (C.c_int, [C.c_void_p, c_uint8_p, c_double_p]),
    "bt_forward_map": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p, c_double_p]),
    "bt_pmala_site": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_int64,
                                c_float_p, c_float_p]),
    "bt_mtm_site": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_int64, c_float_p, c_float_p, c_float_p]),
    "bt_mtm_finish": (C.c_int, [C.c_void_p, C.c_double]),
    "bt_begin_iteration": (C.c_int, [C.c_void_p]),
    "bt_get_step_stats": (C.c_int, [C.c_void_p, c_double_p, c_double_p]),
    "bt_reduce_step_stats": (C.c_int, [C.c_void_p, c_uint8_p, c_double_p]),
    "bt_debug_philox_pmala": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64, C.c_int64, c_float_p, c_float_p]),
    "bt_debug_philox_mtm": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_int64, c_float_p, c_float_p,
                                      c_float_p]),
    "bt_model_check_reset": (C.c_int, [C.c_void_p]),
    "bt_model_check_update": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int64, c_float_p, c_float_p]),
    "bt_model_check_get": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p, C.c_int]),
    "bt_spatial_prior_unweighted": (C.c_int, [C.c_void_p, c_uint8_p, c_double_p]),
    "bt_chain_reserve": (C.c_int, [C.c_void_p, C.c_int64]),
    "bt_chain_push": (C.c_int, [C.c_void_p]),
    "bt_chain_length": (C.c_int64, [C.c_void_p]),
    "bt_chain_get": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, c_double_p]),
    "bt_chain_stats": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int, c_double_p, c_double_p, c_double_p,
                                 c_double_p, c_double_p, c_double_p]),
    "bt_model_check_snapshot_clppd": (C.c_int, [C.c_void_p]),
    "bt_model_check_update_pvalues": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int64, c_float_p, c_float_p]),
    "bt_optim_begin": (C.c_int, [C.c_void_p]),
    "bt_optim_pmala_propose": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_uint64, C.c_int64, c_float_p]),
    "bt_optim_update_v": (C.c_int, [C.c_void_p, C.c_double]),
    "bt_optim_end": (C.c_int, [C.c_void_p, C.c_int]),
    "bt_mtm_optim_weights": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_int64, c_float_p, c_uint8_p,
                                       c_double_p]),
    "bt_mtm_optim_select": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "bt_mlp_kernel_variant": (C.c_int, [C.c_void_p]),
    "bt_kernel_launches": (C.c_int64, [C.c_void_p]),
    "bt_kernel_profile": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p, c_double_p, c_int64_p]),
    "bt_reserve_scratch": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64]),
    "bt_mlp_profile": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p, c_int64_p, c_int64_p]),
}


class BeetrootsB200Error(RuntimeError):
    pass


def load():
    """Load the CUDA library; raise (never fall back) if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BeetrootsB200Error(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "beetroots_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().bt_last_error().decode("utf-8", "replace")
        raise BeetrootsB200Error(f"{what} failed (code {rc}): {msg}")


def dptr(a, ctype=C.c_double):
    """Pointer to a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(ctype))


def as_f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def as_f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def as_i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)
