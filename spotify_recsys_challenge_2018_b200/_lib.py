"""ctypes binding of libdae_b200.so (include/dae_b200.h).  Fails loudly when the library is absent."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdae_b200.so")


class DaeConfig(C.Structure):
    _fields_ = [
        ("n_input", C.c_int32), ("n_tracks", C.c_int32), ("n_hidden", C.c_int32), ("max_batch", C.c_int32),
        ("tied", C.c_int32), ("lr", C.c_float), ("reg_lambda", C.c_float), ("seed", C.c_uint64),
        ("device", C.c_int32), ("trainable", C.c_int32), ("stream", C.c_void_p),
        ("world", C.c_int32), ("rank", C.c_int32),
    ]


class DaeTitleConfig(C.Structure):
    _fields_ = [
        ("charsize", C.c_int32), ("strmaxlen", C.c_int32), ("char_emb", C.c_int32), ("filter_num", C.c_int32),
        ("n_filter_sizes", C.c_int32), ("filter_size", C.c_int32 * 8), ("lr", C.c_float), ("trainable", C.c_int32),
    ]


class DaeError(RuntimeError):
    pass


_P = C.c_void_p
_I32, _I64, _F = C.c_int32, C.c_int64, C.c_float

# name -> (restype, argtypes); every symbol include/dae_b200.h declares
SIGNATURES = {
    "dae_abi_version": (_I32, []),
    "dae_last_error": (C.c_char_p, []),
    "dae_model_create": (_I32, [C.POINTER(DaeConfig), C.POINTER(_P)]),
    "dae_model_destroy": (None, [_P]),
    "dae_model_init_xavier": (_I32, [_P, C.c_uint64]),
    "dae_model_set_params": (_I32, [_P, _P, _P, _P, _P]),
    "dae_model_get_params": (_I32, [_P, _P, _P, _P, _P]),
    "dae_model_get_adam_state": (_I32, [_P, _P, _P, _P, _P, C.POINTER(_I64)]),
    "dae_model_train_step": (_I32, [_P, _P, _P, _I64, _P, _P, _I64, _I32, _F, _F, C.POINTER(_F)]),
    "dae_model_train_step_async": (_I32, [_P, _P, _P, _I64, _P, _P, _I64, _I32, _F, _F, C.POINTER(_F), C.POINTER(_I32)]),
    "dae_model_train_flush": (_I32, [_P, C.POINTER(_F), C.POINTER(_I32)]),
    "dae_model_predict": (_I32, [_P, _P, _P, _I64, _I32, _I32, _P]),
    "dae_model_recommend": (_I32, [_P, _P, _P, _I64, _I32, _P, _P, _I32, _P, _P]),
    "dae_model_recommend_range": (_I32, [_P, _P, _P, _I64, _I32, _P, _P, _I32, _I32, _I32, _P, _P]),
    "dae_model_evaluate": (_I32, [_P, _P, _P, _I64, _I32, _P, _P, _P, _P, _I32, _P]),
    "dae_model_stage_batch": (_I32, [_P, _I32, _P, _P, _I64, _P, _P, _I64, _I32]),
    "dae_model_restage": (_I32, [_P, _I32]),
    "dae_model_backward_staged": (_I32, [_P, _I32, _F, _F, _I32, _I32]),
    "dae_model_apply_adam": (_I32, [_P]),
    "dae_model_train_step_staged": (_I32, [_P, _I32, _F, _F]),
    "dae_model_sync_cost": (_I32, [_P, C.POINTER(_F)]),
    "dae_model_ipc_handle": (_I32, [_P, _P]),
    "dae_model_attach_ipc": (_I32, [_P, _P, _I32]),
    "dae_model_attach_local": (_I32, [_P, C.POINTER(_P), _I32]),
    "dae_model_arena_bytes": (_I32, [_P, C.POINTER(_I64)]),
    "dae_model_set_debug": (_I32, [_P, _I32]),
    "dae_model_buffer": (_I32, [_P, C.c_char_p, C.POINTER(_P), C.POINTER(_I64), C.POINTER(_I32)]),
    "dae_model_launch_count": (_I64, [_P]),
    "dae_model_set_profiling": (_I32, [_P, _I32]),
    "dae_model_phase_count": (_I32, []),
    "dae_model_phase_name": (C.c_char_p, [_I32]),
    "dae_model_phase_time": (_I32, [_P, _I32, C.POINTER(C.c_double), C.POINTER(_I64)]),
    "dae_title_create": (_I32, [_P, C.POINTER(DaeTitleConfig), C.POINTER(_P)]),
    "dae_title_destroy": (None, [_P]),
    "dae_title_init": (_I32, [_P, C.c_uint64]),
    "dae_title_param_count": (_I32, [_P]),
    "dae_title_param_size": (_I32, [_P, _I32, C.POINTER(_I64)]),
    "dae_title_set_params": (_I32, [_P, C.POINTER(_P)]),
    "dae_title_get_params": (_I32, [_P, C.POINTER(_P)]),
    "dae_title_train_step": (_I32, [_P, _P, _P, _I64, _P, _P, _I64, _P, _P, _I32, _F, _F, _F, C.POINTER(_F)]),
    "dae_title_train_step_async": (_I32, [_P, _P, _P, _I64, _P, _P, _I64, _P, _P, _I32, _F, _F, _F, C.POINTER(_F), C.POINTER(_I32)]),
    "dae_title_train_flush": (_I32, [_P, C.POINTER(_F), C.POINTER(_I32)]),
    "dae_title_predict": (_I32, [_P, _P, _P, _I64, _P, _P, _I32, _I32, _P]),
    "dae_title_recommend": (_I32, [_P, _P, _P, _I64, _P, _P, _I32, _P, _P, _I32, _P, _P]),
    "dae_title_evaluate": (_I32, [_P, _P, _P, _I64, _P, _P, _I32, _P, _P, _P, _P, _I32, _P]),
    "dae_title_launch_count": (_I64, [_P]),
    "dae_title_buffer": (_I32, [_P, C.c_char_p, C.POINTER(_P), C.POINTER(_I64), C.POINTER(_I32)]),
    "dae_exchange_create": (_I32, [_I32, _I32, _I32, _I32, _I32, C.POINTER(_P)]),
    "dae_exchange_destroy": (None, [_P]),
    "dae_exchange_ipc_handle": (_I32, [_P, _P]),
    "dae_exchange_attach_ipc": (_I32, [_P, _P, _I32]),
    "dae_exchange_attach_local": (_I32, [_P, C.POINTER(_P), _I32]),
    "dae_exchange_merge_topk": (_I32, [_P, _P, _P, _I32, _I32, _P, _P, _P]),
    "dae_exchange_merge_topk_rows": (_I32, [_P, _P, _P, _I32, _I32, C.POINTER(_I32), C.POINTER(_I32), _P, _P, _P]),
    "dae_model_set_threshold_exchange": (_I32, [_P, _P]),
    "dae_exchange_set_profiling": (_I32, [_P, _I32]),
    "dae_exchange_phase_ms": (_I32, [_P, C.POINTER(C.c_float)]),
    "dae_exchange_launch_count": (_I64, [_P]),
    "dae_topk_device": (_I32, [_P, _I64, _I32, _I32, _I32, _P, _P, _I32, _P, _P, _P]),
    "dae_metrics_device": (_I32, [_P, _I64, _I32, _I32, _P, _P, _P, _P]),
    "dae_topk_merge_device": (_I32, [_P, _P, _I32, _I32, _I32, _P, _P, _P]),
    "dae_adam_device": (_I32, [_P, _P, _P, _P, _P, _I64, _F, _F, _F, _F, _P]),
    "dae_coo_to_csr_device": (_I32, [_P, _P, _I64, _I32, _I32, _P, _P, _P, _P, _P]),
    "dae_dh_nsplit": (_I32, [_I32]),
    "dae_gemm_test_device": (_I32, [_I32, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, C.POINTER(_I32), _P]),
}

_lib = None


def load():
    """Load the shared library (once).  No fallback: a missing build is an error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DaeError(
            "libdae_b200.so is not built (%s).  Run `python -c 'import __graft_entry__ as g; g.build()'` or "
            "`make -C spotify_recsys_challenge_2018_b200/csrc`.  There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the header and the library diverge
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise DaeError(load().dae_last_error().decode("utf-8", "replace"))
